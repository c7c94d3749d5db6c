"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): the single-process multi-device
index and the one-process-per-GPU NCCL path must return exactly what one GPU returns."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus() -> int:
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_multi_device_flat_index_matches_single_gpu(metric):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import vectordb_retrieval_b200.algorithms as A
    rng = np.random.RandomState(3)
    base = rng.randn(30011, 48).astype(np.float32)
    queries = rng.randn(257, 48).astype(np.float32)
    one = A.get_algorithm_instance("ExactSearch", 48, name="one", metric=metric, device="cuda:0")
    many = A.get_algorithm_instance("ExactSearch", 48, name="many", metric=metric, devices=list(range(_n_gpus())))
    one.build_index(base)
    many.build_index(base)
    d1, i1 = one.batch_search(queries, 50)
    d2, i2 = many.batch_search(queries, 50)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(d1, d2)
    ref = oracle.faiss_flat_search(base, queries, 50, metric)
    res = oracle.compare_topk(ref[0], ref[1], d2, i2, rtol=1e-5, atol=0.0 if metric == "l2" else 1e-4)
    assert res["ok"], res


_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
rank = int(os.environ["RANK"]); dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
import vectordb_retrieval_b200.algorithms as A
rng = np.random.RandomState(3)
base = rng.randn(30011, 48).astype(np.float32); queries = rng.randn(257, 48).astype(np.float32)
for mode, exchange in (("rows", "alltoall"), ("rows", "allgather"), ("queries", "alltoall")):
    # row shards + packed exchange + merge kernel (both exchange plans) / replicated base + query slices
    algo = A.get_algorithm_instance("ExactSearch", 48, name="dist", metric="l2", device=dev, shard=mode, exchange=exchange)
    algo.build_index(base)
    tag = mode if exchange == "alltoall" else mode + "_" + exchange
    for rep in range(2):                      # host path: results land in the ranks' shared pinned block
        d, i = algo.batch_search(queries, 50)
    np.save(os.path.join({out!r}, f"d_{{tag}}{{rank}}.npy"), d); np.save(os.path.join({out!r}, f"i_{{tag}}{{rank}}.npy"), i)
    dd, ii = algo.index.search_device(torch.from_numpy(queries).to(dev), 50)      # device path: result on every rank
    np.save(os.path.join({out!r}, f"d_{{tag}}_dev{{rank}}.npy"), dd.cpu().numpy()); np.save(os.path.join({out!r}, f"i_{{tag}}_dev{{rank}}.npy"), ii.cpu().numpy())
    del algo
# IVF-Flat with replicated centroids and row-sharded lists: same answer as one GPU holding every list
from vectordb_retrieval_b200 import engine, sharded
ivf = sharded.DistributedIVFIndex.from_global(base, 64, "l2", dev, nprobe=8)
qd = torch.from_numpy(queries).to(dev)
d, i = ivf.search(qd.clone(), 50)
np.save(os.path.join({out!r}, f"d_ivf{{rank}}.npy"), d.cpu().numpy()); np.save(os.path.join({out!r}, f"i_ivf{{rank}}.npy"), i.cpu().numpy())
# ... and with every list on every rank, the queries cut into slices (device result, then the host path)
rep = sharded.ReplicatedIVFIndex(base, ivf.shard.centroids, "l2", dev, nprobe=8)
d, i = rep.search(qd.clone(), 50)
np.save(os.path.join({out!r}, f"d_ivfrep{{rank}}.npy"), d.cpu().numpy()); np.save(os.path.join({out!r}, f"i_ivfrep{{rank}}.npy"), i.cpu().numpy())
for _ in range(2):
    dh, ih = rep.search_host(queries, 50)
np.save(os.path.join({out!r}, f"d_ivfrep_host{{rank}}.npy"), dh); np.save(os.path.join({out!r}, f"i_ivfrep_host{{rank}}.npy"), ih)
# the same two layouts behind the plugin API (ApproximateSearch under torchrun), with the centroids of rank 0
for mode in ("queries", "rows"):
    algo = A.get_algorithm_instance("ApproximateSearch", 48, name="ivf", index_type="IVF64,Flat", metric="l2", nprobe=8, device=dev, shard=mode)
    algo.build_index(base)
    da, ia = algo.batch_search(queries, 50)
    one = engine.IVFShard(base, algo.index.centroids, "l2", dev)
    d1, i1 = one.search(qd.clone(), 50, 8, 0, engine.FLT_MAX)
    assert np.array_equal(ia, i1.cpu().numpy()) and np.array_equal(da, d1.cpu().numpy()), mode
    assert type(algo.index._dist).__name__ == ("ReplicatedIVFIndex" if mode == "queries" else "DistributedIVFIndex")
    del algo, one
if rank == 0:
    one = engine.IVFShard(base, ivf.shard.centroids, "l2", dev)
    d1, i1 = one.search(qd.clone(), 50, 8, 0, engine.FLT_MAX)
    np.save(os.path.join({out!r}, "d_ivf_single.npy"), d1.cpu().numpy()); np.save(os.path.join({out!r}, "i_ivf_single.npy"), i1.cpu().numpy())
dist.destroy_process_group()
"""


def test_distributed_flat_index_nccl(tmp_path):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, out=str(tmp_path)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
                    "127.0.0.1", "--master-port", str(port), str(script)], check=True, timeout=600)
    rng = np.random.RandomState(3)
    base = rng.randn(30011, 48).astype(np.float32)
    queries = rng.randn(257, 48).astype(np.float32)
    ref = oracle.faiss_flat_search(base, queries, 50, "l2")
    for mode in ("rows", "rows_allgather", "queries"):
        for r in range(n):
            for path in ("", "_dev"):         # host path (shared pinned block) and device path
                res = oracle.compare_topk(ref[0], ref[1], np.load(tmp_path / f"d_{mode}{path}{r}.npy"),
                                          np.load(tmp_path / f"i_{mode}{path}{r}.npy"), rtol=1e-5)
                assert res["ok"], (mode, path, r, res)
            np.testing.assert_array_equal(np.load(tmp_path / f"i_{mode}{r}.npy"), np.load(tmp_path / f"i_{mode}_dev{r}.npy"))
            np.testing.assert_array_equal(np.load(tmp_path / f"d_{mode}{r}.npy"), np.load(tmp_path / f"d_{mode}_dev{r}.npy"))
        np.testing.assert_array_equal(np.load(tmp_path / f"i_{mode}0.npy"), np.load(tmp_path / f"i_{mode}{n - 1}.npy"))
    np.testing.assert_array_equal(np.load(tmp_path / "i_rows0.npy"), np.load(tmp_path / "i_queries0.npy"))
    np.testing.assert_array_equal(np.load(tmp_path / "i_rows0.npy"), np.load(tmp_path / "i_rows_allgather0.npy"))
    np.testing.assert_array_equal(np.load(tmp_path / "d_rows0.npy"), np.load(tmp_path / "d_rows_allgather0.npy"))
    for r in range(n):        # sharded IVF == single-GPU IVF with the same centroids, on every rank
        np.testing.assert_array_equal(np.load(tmp_path / f"i_ivf{r}.npy"), np.load(tmp_path / "i_ivf_single.npy"))
        np.testing.assert_array_equal(np.load(tmp_path / f"d_ivf{r}.npy"), np.load(tmp_path / "d_ivf_single.npy"))
        for tag in ("ivfrep", "ivfrep_host"):     # replicated lists, query slices: the same again
            np.testing.assert_array_equal(np.load(tmp_path / f"i_{tag}{r}.npy"), np.load(tmp_path / "i_ivf_single.npy"))
            np.testing.assert_array_equal(np.load(tmp_path / f"d_{tag}{r}.npy"), np.load(tmp_path / "d_ivf_single.npy"))
