#!/usr/bin/env python
"""Flat search on a GPU's row share of the SIFT1M shape (1M / 8, / 4, / 2 rows) and on query slices, for the seeding
knobs: search time, main-scan time and re-scanned queries per (rows, queries, sample tiles, margin).  Tuning aid."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vectordb_retrieval_b200 import _lib, engine  # noqa: E402


def main():
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    base = torch.randn((1_000_000, 128), generator=g, device=dev)
    q = torch.randn((10_000, 128), generator=g, device=dev)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    for rows, nq in ((125_000, 10_000), (250_000, 10_000), (500_000, 10_000), (1_000_000, 10_000), (1_000_000, 1_250)):
        shard = engine.FlatShard(base[:rows], "l2", dev)
        for tiles, rank, margin in ((64, 0, 0), (64, 16, 4), (64, 16, 3), (64, 32, 3), (64, 32, 4), (128, 32, 3)):
            lib.vdb_flat_set_seeding(tiles, rank)
            lib.vdb_flat_set_seeding_margin(margin)
            for _ in range(3):
                shard.search(q[:nq], 100)
            redo = ctypes.c_uint64(0)
            lib.vdb_debug_redo_queries(ctypes.byref(redo))
            lib.vdb_flat_timing_enable(1)
            ts = []
            for _ in range(8):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); shard.search(q[:nq], 100); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            buf = (ctypes.c_float * 512)(); n = ctypes.c_int(0)
            lib.vdb_flat_timing_read(buf, 512, ctypes.byref(n))
            lib.vdb_flat_timing_enable(0)
            lib.vdb_debug_redo_queries(ctypes.byref(redo))
            scan = sorted(buf[i] for i in range(n.value))[n.value // 2]
            print(json.dumps({"rows": rows, "nq": nq, "sample_tiles": tiles, "rank": rank, "margin": margin, "search_ms": sorted(ts)[len(ts) // 2],
                              "main_scan_ms": scan, "redo_queries_in_8_searches": int(redo.value)}), flush=True)
        del shard
    lib.vdb_flat_set_seeding(64, 0)
    lib.vdb_flat_set_seeding_margin(0)


if __name__ == "__main__":
    main()
