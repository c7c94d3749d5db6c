#!/usr/bin/env python
"""BASELINE.json configs[2] (GloVe-50 shape: 1.2M x 50 fp32, cosine, 10k queries, k = 100) at 1/2/4/8 GPUs: the exact
flat search and IVF-Flat (nlist 4096), both ROW-sharded (every rank holds the rows [lo, hi) - flat operands, or the
inverted lists cut by row range - plus the packed top-k exchange and the merge kernel); ``--layout replicated`` times
IVF-Flat with every list on every rank and the queries cut into slices instead.  Strong scaling: the base is fixed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/scale_c3.py

Parity inside the run (outside the timed steps): rank 0 also holds the WHOLE base as one IVF shard with the same
centroids; the distributed result has to equal it bit for bit (ids and distances) at every nprobe, and the flat result
has to equal the one-shard flat search.  One JSON object per line on rank 0."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vectordb_retrieval_b200 import engine, sharded  # noqa: E402
from vectordb_retrieval_b200.harness.dataset import Dataset  # noqa: E402
from vectordb_retrieval_b200.harness.metrics import recall_at_k  # noqa: E402


def timed(fn, dev, world, reps):
    for _ in range(2):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for e0, e1 in ev:
        e0.record(); out = fn(); e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), out


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_200_000)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--nlist", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--layout", choices=["rows", "replicated", "both"], default="rows")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    d, k = 50, 100
    ds = Dataset("glove50_shape", options={"train_size": args.n, "test_size": args.nq, "ground_truth": "skip", "seed": 42})
    ds._clustered(d, args.n, args.nq, 64, 0.3)                       # seeded on the host: the same base on every rank
    base, queries = ds.train_vectors, ds.test_vectors
    q = torch.from_numpy(queries).to(dev)

    def emit(obj):
        if rank == 0:
            print(json.dumps(dict(obj, n_gpus=world, workload=f"{args.n} x {d} fp32 cosine, {args.nq} queries, k={k}")), flush=True)

    if args.layout in ("replicated", "both"):
        # every list on every rank, the queries cut into slices: nothing is repeated per rank, one allgather of the result blocks
        rep = sharded.ReplicatedIVFIndex.from_global(base, args.nlist, "cosine", dev)
        whole = rep.shard                                                   # IS the one-GPU index
        _, pad = sharded._conventions(engine, "cosine", 0, None)
        flat1 = engine.FlatShard(base, "cosine", dev)
        gt = flat1.search(q, k, 0, pad)[1].cpu().numpy()
        del flat1
        for nprobe in (8, 32, 128):
            rep.nprobe = nprobe
            ms, (D, I) = timed(lambda: rep.search(q, k), dev, world, args.steps)
            D1, I1 = whole.search(q, k, nprobe, 0, pad)
            emit({"algo": "ivf_flat_replicated", "nlist": args.nlist, "nprobe": nprobe, "ms_per_step": ms, "qps": args.nq / ms * 1e3,
                  "recall@100": recall_at_k(gt, I.cpu().numpy(), 100), "equals_one_gpu_index": bool(torch.equal(I1, I) and torch.equal(D1, D))})
        del rep, whole
        if args.layout == "replicated":
            if world > 1:
                dist.destroy_process_group()
            return 0

    flat = sharded.DistributedFlatIndex.from_global(base, "cosine", dev)
    ms, (D, I) = timed(lambda: flat.search(q, k), dev, world, args.steps)
    gt = I.cpu().numpy()
    line = {"algo": "exact_flat_rows", "ms_per_step": ms, "qps": args.nq / ms * 1e3,
            "tf32_pipe_tflops_per_gpu": 3 * 2.0 * args.nq * args.n * 64 / world / (ms * 1e-3) / 1e12}
    if rank == 0 and world > 1:                                          # the one-shard search of the same base
        one_shard = engine.FlatShard(base, "cosine", dev)
        descending, pad = sharded._conventions(engine, "cosine", 0, None)
        D1, I1 = one_shard.search(q, k, 0, pad)
        line["id_mismatch_vs_one_gpu"] = int((I1 != I).sum().item())
        line["max_abs_diff_vs_one_gpu"] = float((D1 - D).abs().max().item())
        del one_shard
    emit(line)
    del flat

    ivf = sharded.DistributedIVFIndex.from_global(base, args.nlist, "cosine", dev)
    whole = engine.IVFShard(base, ivf.shard.centroids, "cosine", dev) if rank == 0 else None     # the one-GPU index, same centroids
    for nprobe in (8, 32, 128):
        ivf.nprobe = nprobe
        ms, (D, I) = timed(lambda: ivf.search(q, k), dev, world, args.steps)
        line = {"algo": "ivf_flat_rows", "nlist": args.nlist, "nprobe": nprobe, "ms_per_step": ms, "qps": args.nq / ms * 1e3,
                "recall@100": recall_at_k(gt, I.cpu().numpy(), 100)}
        if rank == 0:
            _, pad = sharded._conventions(engine, "cosine", 0, None)
            D1, I1 = whole.search(q, k, nprobe, 0, pad)
            line["equals_one_gpu_index"] = bool(torch.equal(I1, I) and torch.equal(D1, D))
            line["id_mismatch_vs_one_gpu"] = int((I1 != I).sum().item())
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
