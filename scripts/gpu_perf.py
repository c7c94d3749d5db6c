#!/usr/bin/env python
"""Quick device-side timing of the flat scan (inputs resident in HBM, CUDA events).
    python scripts/gpu_perf.py [impl] [n] [d] [nq] [k] [metric]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vectordb_retrieval_b200 import _lib, engine  # noqa: E402

impl = sys.argv[1] if len(sys.argv) > 1 else "tcgen05"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 128
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 10_000
k = int(sys.argv[5]) if len(sys.argv) > 5 else 100
metric = sys.argv[6] if len(sys.argv) > 6 else "l2"
reps = int(os.environ.get("REPS", "5"))
g = torch.Generator(device="cuda").manual_seed(42)
base = torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32)
q = torch.randn((nq, d), generator=g, device="cuda", dtype=torch.float32)
shard = engine.FlatShard(base, metric, "cuda")
del base
code = _lib.IMPL_NAMES[impl]
dbg = int(os.environ.get("VDB_DBG", "0"))
if dbg:
    _lib.load().vdb_set_debug_mode(dbg)
if os.environ.get("VDB_SEED"):                      # "sample_tiles,rank" of the seeding pre-pass (0 tiles = off)
    _t, _r = (int(v) for v in os.environ["VDB_SEED"].split(","))
    assert _lib.load().vdb_flat_set_seeding(_t, _r) == 0
for _ in range(2):
    D, I = shard.search(q.clone(), k, 0, 3.4e38, code)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    qq = q.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    D, I = shard.search(qq, k, 0, 3.4e38, code)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
if dbg >= 8:
    import ctypes
    buf = (ctypes.c_uint64 * 8)()
    _lib.load().vdb_debug_read_prof(buf)
    _ = shard.search(q.clone(), k, 0, 3.4e38, code)
    torch.cuda.synchronize()
    _lib.load().vdb_debug_read_prof(buf)
    tot, wait, hitcyc, app, hits0, hitcyc0, warps, hits = [int(buf[i]) for i in range(8)]
    w = max(warps, 1)
    print(f"  epilogue warps {warps}: {tot / w:,.0f} cycles each, {100 * wait / max(tot, 1):.1f}% waiting for an accumulator, "
          f"{100 * hitcyc / max(tot, 1):.1f}% in the append path ({hits / w:,.0f} chunks with a hit per warp, "
          f"{hitcyc / max(hits, 1):,.0f} cycles each; first-generation items: {hits0 / w:,.0f} chunks, "
          f"{hitcyc0 / max(hits0, 1):,.0f} cycles each); {app:,} candidates appended ({app / nq:,.0f} per query)", flush=True)
import ctypes as _ct
_redo = _ct.c_uint64(0)
_lib.load().vdb_debug_redo_queries(_ct.byref(_redo))
if _redo.value:
    print(f"  queries re-scanned over all calls: {_redo.value}")
ms = float(np.median(ts))
flops = 2.0 * nq * n * d
print(f"[{impl} dbg={dbg}] n={n} d={d} nq={nq} k={k} {metric}: median {ms:.3f} ms (min {min(ts):.3f})  "
      f"QPS {nq / ms * 1e3:,.0f}  fp32-equiv {flops / ms / 1e9:.1f} TFLOP/s  tf32-pipe(3x) {3 * flops / ms / 1e9:.1f} TFLOP/s", flush=True)
