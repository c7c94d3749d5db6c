#!/usr/bin/env python
"""C3 shape (1.2M x 50 cosine, nlist 4096, 10k queries, k = 100): three IVF-Flat searches at one nprobe, for an ncu capture of
`ivf_scan_kernel` at that launch shape (`-k regex:ivf_scan_kernel --launch-skip 2 --launch-count 1`).  Prints the event-timed
search and the scanned bytes so the capture can be set against a plain run.    python scripts/profile_ivf.py --nprobe 8"""
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
    ap.add_argument("--nprobe", type=int, default=8)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n, d, nq, nlist, k = 1_200_000, 50, 10_000, 4096, 100
    ds = Dataset("glove50_shape", options={"train_size": n, "test_size": nq, "ground_truth": "skip", "seed": 42})
    ds._clustered(d, n, nq, 64, 0.3)
    cent = engine.kmeans_train(ds.train_vectors, nlist, "cosine", dev, niter=10)
    ivf = engine.IVFShard(ds.train_vectors, cent, "cosine", dev)
    q = torch.from_numpy(ds.test_vectors).to(dev)
    scanned = torch.zeros(1, dtype=torch.int64, device=dev)
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ivf.search(q, k, args.nprobe, 0, -engine.FLT_MAX, scanned if rep == 0 else None); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    rows = int(scanned.item())
    print(json.dumps({"nprobe": args.nprobe, "search_ms": min(ts), "scanned_rows": rows, "algorithmic_gb": rows * d * 4 / 1e9,
                      "list_bytes_gb_incl_padding": rows * ivf.d4 * 16 / 1e9}))


if __name__ == "__main__":
    main()
