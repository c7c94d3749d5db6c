#!/usr/bin/env python
"""BASELINE.json configs[4]: synthetic 100M x 128 fp32 base, 10k queries, exact L2 top-100, at
1/2/4/8 GPUs (row-sharded, one packed NCCL exchange of the local top-k lists + merge kernel).  Strong scaling:
the base is fixed.
With ``--rows 8800000 --dim 768 --metric ip`` it is configs[3] (MS MARCO pre-embedded shape, 8 GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/scale_100m.py [--rows 100000000]

The base never exists on the host: every rank generates its rows on the device in 65 536-row blocks
seeded by block number (identical data for every GPU count).  Parity at this size (outside the timed steps):
sorted distances, ids in range, and - for 16 sampled queries - an independent brute force in torch fp64:
every rank regenerates its blocks, scores them against the sampled queries with an fp64 matmul, keeps its
best 100, rank 0 merges the ranks' lists by (distance, id) and compares them with the distributed result
(ids equal except where fp64 distances tie within 1e-5 relative, distances within 1e-5 relative)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vectordb_retrieval_b200 import engine, sharded  # noqa: E402

NQ, TOPK, BLK = 10_000, 100, 65536


def brute_force_check(rows, lo, hi, q, D, I, args, rank, world, dev):
    """Independent fp64 brute force for a query sample (see the module docstring); returns the parity dict on rank 0."""
    import numpy as np
    pick = torch.from_numpy(np.sort(np.random.RandomState(11).choice(NQ, args.check_queries, replace=False))).to(dev)
    q64 = q[pick].to(torch.float64)
    qn = (q64 * q64).sum(dim=1, keepdim=True)
    best_v = torch.full((pick.numel(), 0), 0.0, dtype=torch.float64, device=dev)
    best_i = torch.full((pick.numel(), 0), 0, dtype=torch.int64, device=dev)
    step = 8 * BLK
    for s in range(0, hi - lo, step):
        blk = rows[slice(s, min(s + step, hi - lo))].to(torch.float64)
        ip = q64 @ blk.T
        key = (qn + (blk * blk).sum(dim=1)[None, :] - 2.0 * ip) if args.metric == "l2" else -ip
        kk = min(TOPK, key.shape[1])
        v, i = torch.topk(key, kk, dim=1, largest=False)
        best_v = torch.cat([best_v, v], dim=1)
        best_i = torch.cat([best_i, i + lo + s], dim=1)
        if best_v.shape[1] > 4 * TOPK:
            v, o = torch.topk(best_v, TOPK, dim=1, largest=False)
            best_v, best_i = v, torch.gather(best_i, 1, o)
    v, o = torch.topk(best_v, min(TOPK, best_v.shape[1]), dim=1, largest=False)
    best_v, best_i = v.contiguous(), torch.gather(best_i, 1, o).contiguous()
    if world > 1:
        gv = [torch.empty_like(best_v) for _ in range(world)]
        gi = [torch.empty_like(best_i) for _ in range(world)]
        dist.all_gather(gv, best_v)
        dist.all_gather(gi, best_i)
        best_v, best_i = torch.cat(gv, dim=1), torch.cat(gi, dim=1)
    if rank != 0:
        return None
    ref_v, ref_i = best_v.cpu().numpy(), best_i.cpu().numpy()
    order = np.lexsort((ref_i, ref_v), axis=1)[:, :TOPK]
    ref_v, ref_i = np.take_along_axis(ref_v, order, axis=1), np.take_along_axis(ref_i, order, axis=1)
    got_d = D[pick].cpu().numpy().astype(np.float64)
    got_i = I[pick].cpu().numpy()
    ref_d = np.maximum(ref_v, 0.0) if args.metric == "l2" else -ref_v
    scale = np.maximum(np.abs(ref_d), 1e-30) if args.metric == "l2" else np.abs(ref_d) + float(q64.norm(dim=1).max()) * 12.0
    rel = np.abs(got_d - ref_d) / scale
    id_exact = int((got_i == ref_i).sum())
    mismatch = 0
    for r in range(ref_i.shape[0]):                     # ids may differ only inside fp64 near-ties (1e-5 relative)
        for p in np.nonzero(got_i[r] != ref_i[r])[0]:
            if rel[r, p] > 1e-5 or got_i[r, p] not in set(ref_i[r].tolist()) and abs(got_d[r, p] - ref_d[r, -1]) > 1e-5 * scale[r, -1]:
                mismatch += 1
    return {"queries": int(pick.numel()), "checker": "torch fp64 brute force over the regenerated blocks, all ranks", "rtol": 1e-5,
            "ok": bool(mismatch == 0 and float(rel.max()) <= 1e-5), "id_exact": id_exact, "id_total": int(ref_i.size),
            "id_mismatch": int(mismatch), "max_rel_err": float(rel.max())}


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--metric", choices=["l2", "ip"], default="l2")
    ap.add_argument("--exchange", choices=["alltoall", "allgather"], default="alltoall")
    ap.add_argument("--check-queries", type=int, default=16)
    args = ap.parse_args()
    DIM = args.dim
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    plan = sharded.ShardPlan(args.rows, world)
    lo, hi = plan.start(rank), plan.stop(rank)

    class Rows:                      # rows [lo, hi) materialised block-wise on the device by FlatShard's upload loop
        shape = (hi - lo, DIM)

        def __getitem__(self, sl):
            s, e = lo + sl.start, lo + min(sl.stop, hi - lo)
            out = torch.empty((e - s, DIM), dtype=torch.float32, device=dev)
            for b in range(s // BLK, (e + BLK - 1) // BLK):
                g = torch.Generator(device=dev).manual_seed(42_000 + b)
                blk = torch.randn((BLK, DIM), generator=g, device=dev, dtype=torch.float32)
                a, z = max(s, b * BLK), min(e, (b + 1) * BLK)
                out[a - s:z - s] = blk[a - b * BLK:z - b * BLK]
            return out

    t0 = time.time()
    index = sharded.DistributedFlatIndex(Rows(), args.metric, dev, id_offset=lo, exchange=args.exchange)
    torch.cuda.synchronize(dev)
    build_s = time.time() - t0
    q = torch.randn((NQ, DIM), generator=torch.Generator(device=dev).manual_seed(4242), device=dev)
    for _ in range(2):
        D, I = index.search(q, TOPK)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for e0, e1 in ev:
        e0.record(); D, I = index.search(q, TOPK); e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ordered = (D[:, 1:] >= D[:, :-1]) if args.metric == "l2" else (D[:, 1:] <= D[:, :-1])      # raw inner products descend
    ok = bool(ordered.all().item()) and int(I.min()) >= 0 and int(I.max()) < args.rows
    D, I = D.clone(), I.clone()
    parity = brute_force_check(Rows(), lo, hi, q, D, I, args, rank, world, dev) if args.check_queries > 0 else None
    if rank == 0:
        flops = 2.0 * NQ * args.rows * DIM
        print(json.dumps({"workload": f"{args.rows} x {DIM} fp32, {NQ} queries, exact {args.metric} top-{TOPK}", "n_gpus": world,
                          "ms_per_step": float(ms.item()), "qps": NQ / (float(ms.item()) * 1e-3),
                          "tf32_pipe_tflops_per_gpu": 3 * flops / world / (float(ms.item()) * 1e-3) / 1e12,
                          "build_s": build_s, "operand_gb_per_gpu": index.memory_bytes() / 1e9, "exchange": args.exchange,
                          "sorted_and_in_range": ok, "parity": parity}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
