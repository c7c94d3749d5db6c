#!/usr/bin/env python
"""IVF list scan on the C3 shape (1.2M x 50 cosine, nlist 4096, 10k queries): scan-kernel time per nprobe for 1 / 2 / 4 / 8
warps per query (VDB_IVF_WPQ override) - the data behind the warps-per-query rule.  Tuning aid."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vectordb_retrieval_b200 import _lib, engine  # noqa: E402
from vectordb_retrieval_b200.harness.dataset import Dataset  # noqa: E402


def main():
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    n, d, nq, nlist, k = 1_200_000, 50, 10_000, 4096, 100
    ds = Dataset("glove50_shape", options={"train_size": n, "test_size": nq, "ground_truth": "skip", "seed": 42})
    ds._clustered(d, n, nq, 64, 0.3)
    cent = engine.kmeans_train(ds.train_vectors, nlist, "cosine", dev, niter=10)
    ivf = engine.IVFShard(ds.train_vectors, cent, "cosine", dev)
    q = engine.normalized_rows(torch.from_numpy(ds.test_vectors).to(dev))
    out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    for nprobe in (1, 2, 4, 8, 16, 32, 64):
        _, probes = ivf.quantizer.search(q, nprobe)
        scanned = torch.zeros(1, dtype=torch.int64, device=dev)
        row = {"nprobe": nprobe}
        for w in (1, 2, 4, 8):
            os.environ["VDB_IVF_WPQ"] = str(w)
            ts = []
            for rep in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(lib.vdb_ivf_scan_topk_ex(1, ivf.list_vecs.data_ptr(), ivf.list_ids.data_ptr(), ivf.blk_off.data_ptr(), nlist, d,
                                                    probes.data_ptr(), nprobe, q.data_ptr(), q.stride(0), nq, k, 0, -engine.FLT_MAX, 0,
                                                    out_d.data_ptr(), out_i.data_ptr(), scanned.data_ptr() if rep == 0 and w == 1 else 0,
                                                    1, torch.cuda.current_stream().cuda_stream), "scan")
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            row[f"ms_w{w}"] = round(sorted(ts[1:])[2], 4)
        rows = int(scanned.item())
        best = min(row[f"ms_w{w}"] for w in (1, 2, 4, 8))
        row.update(scanned_rows_per_query=rows / nq, best_frac_of_hbm=rows * d * 4 / (best * 1e-3) / 1e9 / 6550.1)
        print(json.dumps(row), flush=True)
    os.environ.pop("VDB_IVF_WPQ", None)


if __name__ == "__main__":
    main()
