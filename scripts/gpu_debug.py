#!/usr/bin/env python
"""Bring-up probe for the scan kernels: dense key matrix vs fp64 numpy, per implementation.
Run each impl in its own process (a device trap kills the CUDA context):
    python scripts/gpu_debug.py tcgen05 [n d nq]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vectordb_retrieval_b200 import _lib, engine  # noqa: E402

impl = sys.argv[1] if len(sys.argv) > 1 else "tcgen05"
n, d, nq = (int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (1000, 128, 200)
rng = np.random.RandomState(0)
base = rng.randn(n, d).astype(np.float32)
q = rng.randn(nq, d).astype(np.float32)
shard = engine.FlatShard(base, "l2", "cuda")
hi = shard.hi.cpu().numpy()[:n, :d]
lo = shard.lo.cpu().numpy()[:n, :d]
print(f"[{impl}] n={n} d={d} nq={nq} split exact: {np.array_equal(hi + lo, base)}  norms ok: "
      f"{np.allclose(shard.norms.cpu().numpy()[:n], (base.astype(np.float64) ** 2).sum(1), rtol=1e-6)}", flush=True)
t0 = time.time()
keys = shard.dense_keys(torch.from_numpy(q).cuda(), _lib.IMPL_NAMES[impl]).cpu().numpy().astype(np.float64)
ref = (base.astype(np.float64) ** 2).sum(1)[None, :] - 2 * q.astype(np.float64) @ base.astype(np.float64).T
err = np.abs(keys - ref)
print(f"[{impl}] dense keys in {time.time() - t0:.2f}s  max abs err {err.max():.3e}  mean {err.mean():.3e}", flush=True)
if err.max() > 1e-3:
    r, c = np.unravel_index(err.argmax(), err.shape)
    print("worst at", r, c, "got", keys[r, c], "ref", ref[r, c])
    print("row0 got", np.round(keys[0, :12], 3))
    print("row0 ref", np.round(ref[0, :12], 3))
    bad_rows = np.nonzero(err.max(axis=1) > 1e-3)[0]
    bad_cols = np.nonzero(err.max(axis=0) > 1e-3)[0]
    print("bad rows", bad_rows[:20], len(bad_rows), "bad cols", bad_cols[:20], len(bad_cols))
    # is a got-row a permutation / other row of ref?
    for rr in (0, 1, 33, 129):
        if rr < nq:
            best = np.argmin(np.abs(ref[:, :64] - keys[rr, :64][None, :]).sum(1))
            print(f"  got row {rr} best matches ref row {best}")
    sys.exit(1)
for k in (10, 100):
    D, I = shard.search(torch.from_numpy(q).cuda(), k, 0, 3.4e38, _lib.IMPL_NAMES[impl])
    torch.cuda.synchronize()
    order = np.argsort(ref, axis=1, kind="stable")[:, :k]
    same = (I.cpu().numpy() == order).mean()
    print(f"[{impl}] top-{k} id agreement with fp64 argsort: {same:.6f}", flush=True)
print(f"[{impl}] OK")
