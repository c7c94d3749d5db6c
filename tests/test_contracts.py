"""Repository contracts the judge checks mechanically: the product package never touches the
oracle, the bench's reference arm prints the agreed JSON line, graft entry points exist."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vectordb_retrieval_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(base, f), encoding="utf-8").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "oracle." in text.replace("oracle.py", ""):
                    offenders.append(os.path.relpath(os.path.join(base, f), ROOT))
    assert not offenders, f"product files reference the oracle: {offenders}"


def test_graft_entry_points_exist_and_build_runs():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    assert callable(entry.build) and callable(entry.smoke)
    entry.build()          # nvcc cross-compiles without a GPU; no-op when the library is current


def test_bench_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and d["steps"] == 1
