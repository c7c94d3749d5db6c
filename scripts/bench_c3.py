#!/usr/bin/env python
"""BASELINE.json configs[2]: GloVe-50-shape synthetic 1.2M x 50, cosine; IVF-Flat nlist=4096 with an
nprobe sweep next to the LSH (256-bit sign codes) candidate + rerank pipeline.  HBM-bound kernels:
    IVF list scan    algorithmic bytes = sum of scanned list rows * d * 4      
    LSH rerank       algorithmic bytes = nq * C * d * 4
reported as GB/s against the measured copy bandwidth (MEASURED_PEAKS.json), with recall@100 against
the exact GPU search.  One JSON object per line on stdout.   python scripts/bench_c3.py [--n 1200000]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vectordb_retrieval_b200 import engine, indexes  # noqa: E402
from vectordb_retrieval_b200.harness.dataset import Dataset  # noqa: E402
from vectordb_retrieval_b200.harness.metrics import recall_at_k  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_200_000)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--nlist", type=int, default=4096)
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = float(peaks["hbm_gbs"])
    d, k = 50, 100
    ds = Dataset("glove50_shape", options={"train_size": args.n, "test_size": args.nq, "ground_truth": "skip", "seed": 42})
    ds._clustered(d, args.n, args.nq, 64, 0.3)
    base, queries = ds.train_vectors, ds.test_vectors
    dev = torch.device("cuda", 0)
    q_dev = torch.from_numpy(queries).to(dev)

    exact = indexes.GpuIndexFlat(d, "ip", device=dev, normalize=True)
    exact.add(base)
    ms, (D, I) = timed(lambda: exact.search_device(q_dev.clone(), k))
    gt = I.cpu().numpy()
    print(json.dumps({"algo": "exact_flat_cosine", "n": args.n, "d": d, "nq": args.nq, "k": k, "ms": ms, "qps": args.nq / ms * 1e3,
                      "tf32_pipe_tflops": 3 * 2.0 * args.nq * args.n * 64 / (ms * 1e-3) / 1e12, "note": "d padded 50 -> 64"}), flush=True)
    del exact

    t0 = time.time()
    ivf = indexes.GpuIndexIVFFlat(d, args.nlist, "ip", device=dev, normalize=True)
    ivf.train(base)
    t_train = time.time() - t0
    ivf.add(base)
    torch.cuda.synchronize()
    print(json.dumps({"algo": "ivf_build", "nlist": args.nlist, "train_s": t_train, "train_plus_add_s": time.time() - t0,
                      "list_len_mean": args.n / args.nlist, "list_len_max": int(ivf._impl.counts.max().item())}), flush=True)
    shard = ivf._impl
    for nprobe in (1, 2, 4, 8, 16, 32, 64, 128):
        scanned = torch.zeros(1, dtype=torch.int64, device=dev)
        pad = -engine.FLT_MAX
        shard.search(q_dev.clone(), k, nprobe, 0, pad, scanned)
        rows = int(scanned.item())
        ms_total, (D, I) = timed(lambda: shard.search(q_dev.clone(), k, nprobe, 0, pad))
        qn = engine.normalize_rows_(q_dev.clone())
        ms_coarse, _ = timed(lambda: shard.quantizer.search(qn.clone(), nprobe))
        ms_scan = max(ms_total - ms_coarse, 1e-3)
        gbs = rows * d * 4 / (ms_scan * 1e-3) / 1e9
        print(json.dumps({"algo": "ivf_flat", "nprobe": nprobe, "recall@100": recall_at_k(gt, I.cpu().numpy(), 100), "ms_total": ms_total,
                          "ms_coarse": ms_coarse, "ms_list_scan": ms_scan, "qps": args.nq / ms_total * 1e3, "scanned_rows": rows,
                          "algorithmic_gb": rows * d * 4 / 1e9, "achieved_gbs": gbs, "peak_gbs": hbm, "frac": gbs / hbm}), flush=True)
    del ivf, shard

    # IVF<nlist>,SQ8: the same coarse quantiser, 8-bit residual codes (one byte per component in the lists)
    t0 = time.time()
    sq = indexes.GpuIndexIVFSQ8(d, args.nlist, "ip", device=dev, normalize=True)
    sq.train(base)
    sq.add(base)
    torch.cuda.synchronize()
    print(json.dumps({"algo": "ivf_sq8_build", "nlist": args.nlist, "train_plus_add_s": time.time() - t0,
                      "list_bytes_per_row": int(sq._impl.d16 * 16)}), flush=True)
    shard = sq._impl
    for nprobe in (1, 8, 32, 128):
        pad = -engine.FLT_MAX
        ms_total, (D, I) = timed(lambda: shard.search(q_dev.clone(), k, nprobe, 0, pad))
        qn = engine.normalize_rows_(q_dev.clone())
        ms_coarse, _ = timed(lambda: shard.quantizer.search(qn.clone(), nprobe))
        ms_scan = max(ms_total - ms_coarse, 1e-3)
        rows = nprobe * args.n / args.nlist * args.nq            # expected scanned rows (lists are balanced to within the k-means)
        gbs = rows * shard.d16 * 16 / (ms_scan * 1e-3) / 1e9
        print(json.dumps({"algo": "ivf_sq8", "nprobe": nprobe, "recall@100": recall_at_k(gt, I.cpu().numpy(), 100), "ms_total": ms_total,
                          "ms_coarse": ms_coarse, "ms_list_scan": ms_scan, "qps": args.nq / ms_total * 1e3,
                          "approx_code_gb": rows * shard.d16 * 16 / 1e9, "approx_code_gbs": gbs, "peak_gbs": hbm, "frac": gbs / hbm}), flush=True)
    del sq, shard

    # IVF<nlist>,PQ<m> and PQ<m>: m = d one-byte sub-quantisers (the reference's glove50 rows use PQ50), look-up-table scan
    for key in (f"IVF{args.nlist},PQ{d}", f"PQ{d}"):
        t0 = time.time()
        pq = indexes.index_factory(d, key, "ip", device=dev, normalize=True)
        pq.train(base)
        pq.add(base)
        torch.cuda.synchronize()
        print(json.dumps({"algo": "pq_build", "index_key": key, "train_plus_add_s": time.time() - t0,
                          "list_bytes_per_row": int(pq._impl.m16 * 16)}), flush=True)
        # PQ<m>: the table scan, then the flat tensor-pipe scan over the decoded rows (what a query batch gets by default)
        for nprobe, scan in (((8, "table"), (32, "table"), (128, "table")) if key.startswith("IVF") else ((1, "table"), (1, "decoded"))):
            if hasattr(pq, "nprobe"):
                pq.nprobe = nprobe
            pq._impl.decoded_scan = "never" if scan == "table" else "always"
            ms_total, (D, I) = timed(lambda: pq.search_device(q_dev.clone(), k), reps=3)
            rows = (nprobe * args.n / args.nlist if key.startswith("IVF") else args.n) * args.nq
            line = {"algo": "pq", "index_key": key, "nprobe": nprobe, "scan": scan, "recall@100": recall_at_k(gt, I.cpu().numpy(), 100),
                    "ms_total": ms_total, "qps": args.nq / ms_total * 1e3}
            if scan == "table":
                line["table_lookups_per_s"] = rows * d / (ms_total * 1e-3)
                I_table = I
            else:
                line["ids_equal_to_table_scan"] = float((I == I_table).float().mean().item())
                line["decoded_operand_gb"] = pq._impl._flat.memory_bytes() / 1e9
            print(json.dumps(line), flush=True)
        del pq

    lsh = indexes.GpuIndexLSH(d, 256, device=dev)
    lsh.add(base)
    rr = engine.Reranker(base, "cosine", dev)
    for c in (800, 3200, 6400):
        ms_cand, (_, cand) = timed(lambda: lsh.search_device(q_dev.clone(), c), reps=3)
        ms_rr, (D, I) = timed(lambda: rr.search(q_dev.clone(), cand, k, engine._lib.OUT_NEGATE))
        gbs = args.nq * c * d * 4 / (ms_rr * 1e-3) / 1e9
        print(json.dumps({"algo": "faiss_lsh_rerank", "candidates": c, "recall@100": recall_at_k(gt, I.cpu().numpy(), 100),
                          "ms_hamming_topk": ms_cand, "ms_rerank": ms_rr, "qps": args.nq / (ms_cand + ms_rr) * 1e3,
                          "rerank_algorithmic_gb": args.nq * c * d * 4 / 1e9, "rerank_achieved_gbs": gbs, "peak_gbs": hbm,
                          "rerank_frac": gbs / hbm,
                          "hamming_pairs_per_s": args.nq * float(args.n) / (ms_cand * 1e-3)}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
