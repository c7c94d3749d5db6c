#!/usr/bin/env python
"""Hamming top-C on the C3 shape (1.2M x 256-bit codes, 10k queries): CUDA-event time per search for the fp16- and
fp32-accumulator scans.  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vectordb_retrieval_b200 import engine  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(3)
    n, d, nq = 1_200_000, 50, 10_000
    mu = torch.randn((64, d), generator=g, device=dev)
    base = mu[torch.randint(0, 64, (n,), generator=g, device=dev)] + 0.3 * torch.randn((n, d), generator=g, device=dev)
    q = mu[torch.randint(0, 64, (nq,), generator=g, device=dev)] + 0.3 * torch.randn((nq, d), generator=g, device=dev)
    proj = np.random.RandomState(1234).normal(size=(256, d)).astype(np.float32)
    shard = engine.HammingShard(base, proj, dev)
    reps = int(os.environ.get("REPS", "3"))
    for f16 in (True, False):
        shard.tc_accumulate_f16 = f16
        for c in (800, 6400):
            shard.search(q, c)
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); shard.search(q, c); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(json.dumps({"accumulate_f16": f16, "candidates": c, "ms": sorted(ts)[len(ts) // 2]}), flush=True)


if __name__ == "__main__":
    main()
