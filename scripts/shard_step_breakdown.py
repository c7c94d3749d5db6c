#!/usr/bin/env python
"""One GPU's share of a row-sharded SIFT1M-shape step (1M / N rows, all 10k queries, k = 100) plus the merge of N
exchanged lists, launch by launch - run it under `ncu --metrics gpu__time_duration.sum` for the launch list, or plain
for the event-timed totals.  Tuning aid for the fixed per-step costs that limit row-shard scaling at small shards.

    python scripts/shard_step_breakdown.py --parts 8 [--searches 3]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vectordb_retrieval_b200 import engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parts", type=int, default=8)
    ap.add_argument("--searches", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    rows = 1_000_000 // args.parts
    base = torch.randn((rows, 128), generator=g, device=dev)
    q = torch.randn((10_000, 128), generator=g, device=dev)
    shard = engine.FlatShard(base, "l2", dev)
    nq, k, n = q.shape[0], 100, args.parts
    per = (nq + n - 1) // n
    for _ in range(3):
        D, I = shard.search(q, k)
    torch.cuda.synchronize()
    lists_d = D[:per].unsqueeze(0).repeat(n, 1, 1).contiguous()        # [N, nq / N, k]: what the exchange delivers
    lists_i = I[:per].unsqueeze(0).repeat(n, 1, 1).contiguous()
    ts, tm = [], []
    for _ in range(args.searches):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); shard.search(q, k); e[1].record(); engine.merge_topk(lists_d, lists_i); e[2].record()
        torch.cuda.synchronize()
        ts.append(e[0].elapsed_time(e[1])); tm.append(e[1].elapsed_time(e[2]))
    print(json.dumps({"rows": rows, "nq": nq, "search_ms": sorted(ts)[len(ts) // 2], "merge_ms": sorted(tm)[len(tm) // 2],
                      "merge_shape": [n, per, k]}))


if __name__ == "__main__":
    main()
