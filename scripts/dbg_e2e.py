import os, sys, time, torch, numpy as np
import torch.distributed as dist
sys.path.insert(0, os.getcwd())
rank = int(os.environ["RANK"]); dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
from vectordb_retrieval_b200 import engine, sharded
from vectordb_retrieval_b200.indexes import GpuIndexFlat
g = torch.Generator(device=dev).manual_seed(1)
rows = torch.randn((1_000_000, 128), generator=g, device=dev)
index = sharded.ReplicatedFlatIndex(rows, "l2", dev)
q_dev = torch.randn((10000, 128), generator=g, device=dev)
q_host = torch.empty((10000, 128), dtype=torch.float32, pin_memory=True); q_host.copy_(q_dev); qn = q_host.numpy()
gi = GpuIndexFlat(128, "l2", device=dev); gi._impl, gi.ntotal = index, 1_000_000
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
a = t(lambda: index.search(q_dev, 100))
b = t(lambda: index.search_host(qn, 100))
c = t(lambda: gi.search(qn, 100))
d = t(lambda: engine.results_to_host(*index.search(engine.queries_to_device(qn, dev, 128), 100)))
lo, hi, per = index._slice(10000)
e = t(lambda: engine.queries_to_device(qn[lo:hi], dev, 128))
f = t(lambda: engine.queries_to_device(qn, dev, 128))
print(rank, "device %.2f search_host %.2f gi.search %.2f old-path %.2f h2d-slice %.3f h2d-full %.3f" % (a, b, c, d, e, f), torch.from_numpy(qn[lo:hi]).is_pinned(), flush=True)
dist.destroy_process_group()
