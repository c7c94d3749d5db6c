#!/usr/bin/env python
"""Rerank kernel on the C3 shape (1.2M x 50 rows, 10k queries) for C = 800 / 3 200 / 6 400 random candidate ids per query
(the worst case for the gather): time and algorithmic GB/s against the measured copy bandwidth.  Tuning aid."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vectordb_retrieval_b200 import _lib, engine  # noqa: E402


def main():
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    n, d, nq = 1_200_000, 50, 10_000
    base = torch.randn((n, d), generator=g, device=dev)
    q = torch.randn((nq, d), generator=g, device=dev)
    rr = engine.Reranker(base, "cosine", dev)
    peak = 6550.1
    for c in (800, 3200, 6400):
        cand = torch.randint(0, n, (nq, c), generator=g, device=dev, dtype=torch.int64)
        for staged in (0,):
            for _ in range(2):
                rr.search(q, cand, 100, _lib.OUT_NEGATE)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); rr.search(q, cand, 100, _lib.OUT_NEGATE); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[2]
            gbs = nq * c * d * 4 / (ms * 1e-3) / 1e9
            print(json.dumps({"candidates": c, "ms": ms,
                              "algorithmic_gbs": gbs, "frac_of_measured_hbm": gbs / peak}), flush=True)


if __name__ == "__main__":
    main()
