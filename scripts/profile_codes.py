#!/usr/bin/env python
"""C3 shape (1.2M x 50 cosine, nlist 4096, 10k queries, k = 100): three searches of the 8-bit code indexes at one nprobe, for
ncu captures of `ivf_sq8_scan_kernel` / `ivf_pq_scan_kernel` (`-k regex:<kernel> --launch-skip 2 --launch-count 1`).
    python scripts/profile_codes.py --kind sq8|pq [--nprobe 32]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vectordb_retrieval_b200 import engine  # noqa: E402
from vectordb_retrieval_b200.harness.dataset import Dataset  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", choices=["sq8", "pq"], required=True)
    ap.add_argument("--nprobe", type=int, default=32)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n, d, nq, nlist, k = 1_200_000, 50, 10_000, 4096, 100
    ds = Dataset("glove50_shape", options={"train_size": n, "test_size": nq, "ground_truth": "skip", "seed": 42})
    ds._clustered(d, n, nq, 64, 0.3)
    cent = engine.kmeans_train(ds.train_vectors, nlist, "cosine", dev, niter=10)
    if args.kind == "sq8":
        shard = engine.IVFSQ8Shard(ds.train_vectors, cent, "cosine", dev)
        row_bytes = shard.d16 * 16
    else:
        shard = engine.IVFPQShard(ds.train_vectors, cent, d, "cosine", dev, niter=8)
        row_bytes = shard.m16 * 16
    q = torch.from_numpy(ds.test_vectors).to(dev)
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); shard.search(q, k, args.nprobe, 0, -engine.FLT_MAX); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    rows = args.nprobe * n / nlist * nq
    print(json.dumps({"kind": args.kind, "nprobe": args.nprobe, "search_ms": min(ts), "expected_scanned_rows": rows,
                      "code_gb": rows * row_bytes / 1e9}))


if __name__ == "__main__":
    main()
