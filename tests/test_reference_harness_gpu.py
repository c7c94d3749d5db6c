"""The reference's OWN harness, unmodified, on the B200 kernels (north star: "drops into
run_full_benchmark.py unchanged").

``scripts/stage_reference.py`` copies the reference checkout into the git-ignored ``baseline/_ref/`` (it
ships to the GPU box with the snapshot; ``/root/reference`` does not exist there).  Each test starts a fresh
interpreter that calls ``plugin.install()`` and then the reference's entry points:

* ``src.benchmark.runner.BenchmarkRunner`` on the config of the reference's own integration test
  (tests/test_benchmark_runner_modular.py:9-65), same assertions;
* ``scripts/run_full_benchmark.py::main`` (scripts/run_full_benchmark.py:281-320) on
  ``configs/reference_random20k.yaml`` = the published `random` run restricted to the scan + top-k
  algorithms: exact must reach recall 1.0, the Python LSH must reproduce the published
  recall@10 = 0.31914062499999996 / recall@1 = 0.34765625
  (benchmark_results/benchmark_20260305_070532/random/lsh_results.json:44-46) bit for bit, and the
  library's launch counter must have advanced (the kernels, not a fallback, produced the numbers)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

_RUNNER_WORKER = r"""
import json, os, sys
sys.path.insert(0, {root!r})
from vectordb_retrieval_b200 import _lib, plugin
lib = _lib.load()
mods = plugin.import_reference({ref!r})
plugin.install(mods)
from src.benchmark.runner import BenchmarkRunner          # the reference's class, file untouched
import vectordb_retrieval_b200.algorithms as ours
assert mods["modular"].SEARCHER_REGISTRY["LinearSearcher"] is ours.LinearSearcher
before = lib.vdb_launch_count()
runner = BenchmarkRunner({config!r}, output_dir={out!r})
results = runner.run()
json.dump({{"results": results, "output_dir": str(runner.output_dir), "launches": int(lib.vdb_launch_count() - before)}},
          open({dump!r}, "w"), default=str)
"""


def _need_ref():
    if not os.path.isdir(os.path.join(REF, "src", "algorithms")):
        pytest.skip("baseline/_ref is not staged (python scripts/stage_reference.py in the build container)")


def test_reference_benchmark_runner_modular_config_through_plugin(tmp_path):
    _need_ref()
    config = {
        "indexers": {"bf_l2": {"type": "BruteForceIndexer", "metric": "l2"}},
        "searchers": {"linear_l2": {"type": "LinearSearcher", "metric": "l2"}},
        "algorithms": {"bf_linear": {"indexer_ref": "bf_l2", "searcher_ref": "linear_l2", "metric": "l2"}},
        "datasets": [{"name": "random", "metric": "l2", "n_queries": 5, "topk": 5,
                      "dataset_options": {"train_size": 32, "test_size": 6, "ground_truth_k": 5, "dimensions": 3, "seed": 123}}],
        "n_queries": 5, "topk": 5, "repeat": 1, "output_dir": str(tmp_path / "benchmark_outputs"), "seed": 11,
    }
    cfg = tmp_path / "config.yaml"
    cfg.write_text(json.dumps(config))
    dump = tmp_path / "dump.json"
    worker = tmp_path / "worker.py"
    worker.write_text(_RUNNER_WORKER.format(root=ROOT, ref=REF, config=str(cfg), out=str(tmp_path / "fallback_outputs"),
                                            dump=str(dump)))
    subprocess.run([sys.executable, str(worker)], check=True, timeout=600, cwd=str(tmp_path))
    got = json.loads(dump.read_text())
    metrics = got["results"]["random"]["bf_linear"]
    assert metrics["n_train"] == 32 and metrics["n_test"] == 5
    assert metrics["recall@1"] == 1.0                      # 32 x 3 random rows: exact search, no ties
    assert got["launches"] > 0, "the CUDA library launched nothing: the numbers came from somewhere else"
    out_dir = got["output_dir"]
    for name in ("benchmark_summary.md", "one-page-summary.md", "qps_recall_summary.md"):
        assert os.path.exists(os.path.join(out_dir, name)), name
    assert [f for f in os.listdir(out_dir) if f.startswith("qps_recall_") and f.endswith(".svg")]


def test_reference_run_full_benchmark_cli_published_random_run(tmp_path):
    _need_ref()
    out = tmp_path / "results"
    script = os.path.join(ROOT, "scripts", "run_reference_benchmark.py")
    config = os.path.join(ROOT, "configs", "reference_random20k.yaml")
    r = subprocess.run([sys.executable, script, "--config", config, "--output-dir", str(out)], capture_output=True, text=True,
                       timeout=1200, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    runs = sorted(os.listdir(out))
    assert len(runs) == 1 and runs[0].startswith("benchmark_")
    res = json.load(open(os.path.join(out, runs[0], "all_results.json")))["random"]
    assert set(res) == {"exact", "exact_faiss_flat", "ivf_flat", "ivf_sq8", "ivf_pq", "pq", "faiss_lsh", "lsh"}
    for name in ("exact", "exact_faiss_flat"):
        assert res[name]["recall@1"] == 1.0 and res[name]["recall@10"] == 1.0, (name, res[name])
        assert res[name]["n_train"] == 20000 and res[name]["n_test"] == 256 and res[name]["topk"] == 20
    assert res["lsh"]["recall@10"] == 0.31914062499999996       # published value, bit for bit
    assert res["lsh"]["recall@1"] == 0.34765625
    # FAISS-dependent rows: k-means / rotation RNG differ from FAISS's (parity unpinned), so the
    # published 0.4105 / 0.9672 are a neighbourhood, not a pin
    assert 0.30 < res["ivf_flat"]["recall@10"] < 0.55, res["ivf_flat"]["recall@10"]
    assert res["faiss_lsh"]["recall@10"] > 0.88, res["faiss_lsh"]["recall@10"]      # published with FAISS's own rotation: 0.967; here 0.922
    # quantised indexes (published with FAISS's own k-means: ivf_sq8 0.509, ivf_pq 0.509, pq 0.967): neighbourhoods again
    assert 0.40 < res["ivf_sq8"]["recall@10"] < 0.65, res["ivf_sq8"]["recall@10"]
    assert 0.40 < res["ivf_pq"]["recall@10"] < 0.65, res["ivf_pq"]["recall@10"]
    assert res["pq"]["recall@10"] > 0.90, res["pq"]["recall@10"]
    # the exact row's parameters name OUR classes through the reference's describe() plumbing
    assert res["exact"]["parameters"]["searcher"]["type"] == "LinearSearcher"
    assert res["exact"]["qps"] > 0 and res["exact"]["index_memory_mb"] > 0
