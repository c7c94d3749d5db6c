#!/usr/bin/env python
"""Headline benchmark: exact L2 top-100 QPS on the SIFT1M shape (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one query batch: 10 000 queries against the
1 000 000 x 128 fp32 base, k = 100 (query operand split, seeding pre-pass over a strided sample,
tcgen05 main scan with the fused top-k bound, verification + redo launch, exact re-scoring + sort).
At N > 1 (strong scaling) the default layout replicates the 1 GB base and gives every rank nq / N
queries, one NCCL allgather concatenates the result blocks; ``--shard rows`` is the north-star layout:
row shards, NCCL allgather of the local top-k lists, merge kernel.

Printed JSON (rank 0, one line):
  value      QPS with the queries already resident in HBM (device time, CUDA events, max over ranks)
  e2e        QPS through the reference-facing API ``ExactSearch.batch_search`` with HOST query /
             result buffers: pinned-host -> device copy of the queries and device -> host copy of
             (distances, ids) inside the timed region
  roofline   dominant kernel (flat_scan_tc_kernel): tensor-pipe TFLOP/s = 3 * 2*nq*N*d / t
             (3xTF32 issues three MMAs per product) against TF32 peak = measured bf16 / 2
  cpu_baseline  the oracle's FAISS-flat restatement (blocked sgemm + argpartition, all host
             threads) timed on this box's host cores on a bounded sample of the same workload
``--impl reference`` times that CPU port alone (FAISS itself is not installable here; see
DESIGN.md) and prints the same line shape with ``"impl": "reference"``."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_BASE, DIM, NQ, TOPK = 1_000_000, 128, 10_000, 100
WORKLOAD = "sift1m_shape_exact_l2: 1M x 128 fp32 base, 10k queries, k=100 (BASELINE.json configs[1])"
METRIC = "qps_exact_l2_top100_sift1m_shape"
L2_BYTES = 126 * 1024 * 1024


# ------------------------------------------------------------------------------------ helpers
def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": float(p["bf16_tflops"]), "bf16_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "hbm_gbs": float(p["hbm_gbs"]), "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (recipe's clocks line)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.gpu_index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:      # noqa: BLE001 - no nvidia-smi: report nulls
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "power_w_max": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:      # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
                except ValueError:
                    continue
                for nm, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:      # noqa: BLE001
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def _host_data(nq_sample: int):
    import numpy as np
    rng = np.random.default_rng(42)
    base = rng.standard_normal((N_BASE, DIM), dtype=np.float32)
    queries = np.random.default_rng(4242).standard_normal((nq_sample, DIM), dtype=np.float32)
    return base, queries


def _cpu_port_qps(base, queries, reps: int):
    """FAISS-flat restatement on the host cores (oracle port); returns (qps, seconds per pass)."""
    from oracle import oracle
    best = float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        oracle.faiss_flat_search_blas(base, queries, TOPK, "l2")
        best = min(best, time.perf_counter() - t)
    return queries.shape[0] / best, best


def _threads() -> int:
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 0) for i in threadpool_info() if i.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:      # noqa: BLE001
        pass
    return os.cpu_count() or 1


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nq_sample = 200
    base, queries = _host_data(nq_sample)
    for _ in range(max(args.warmup, 1)):
        _cpu_port_qps(base, queries[:50], 1)
    times = []
    for _ in range(args.steps):
        t = time.perf_counter()
        _cpu_port_qps(base, queries, 1)
        times.append(time.perf_counter() - t)
    sec = sum(times) / len(times)
    qps = nq_sample / sec
    sample = f"{nq_sample} of the 10k queries per step against the full 1M x 128 base (QPS is per query, so it carries over)"
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n": N_BASE, "d": DIM, "nq": NQ, "k": TOPK, "metric": "l2",
                   "reference_impl": "oracle port of faiss.IndexFlat.search (blocked fp32 sgemm + argpartition, OpenBLAS); "
                                     "faiss-cpu itself is not installed / installable here"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": _threads(), "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------ our arm
def _device_rows(lo: int, hi: int, dev):
    """Rows [lo, hi) of the synthetic base, generated on the device in 65 536-row blocks seeded by
    block number - the same data for every GPU count."""
    import torch
    blk = 65536
    out = torch.empty((hi - lo, DIM), dtype=torch.float32, device=dev)
    b0 = lo // blk
    for b in range(b0, (hi + blk - 1) // blk):
        g = torch.Generator(device=dev).manual_seed(42_000 + b)
        rows = torch.randn((blk, DIM), generator=g, device=dev, dtype=torch.float32)
        s, e = max(lo, b * blk), min(hi, (b + 1) * blk)
        out[s - lo:e - lo] = rows[s - b * blk:e - b * blk]
    return out


def run_ours(args) -> int:
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        print(json.dumps({"error": "bench.py needs a CUDA device (sm_100a); there is no CPU fallback"}))
        return 1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries ONE JSON line: anything libraries print meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from vectordb_retrieval_b200 import _lib, engine, sharded
    from vectordb_retrieval_b200.algorithms import ExactSearch
    from vectordb_retrieval_b200.indexes import GpuIndexFlat
    lib = _lib.load()

    mode = sharded.choose_sharding(N_BASE, DIM, world, args.shard) if world > 1 else "none"
    if mode == "queries":        # small base: replicate it, every rank searches nq / world queries
        lo, hi = 0, N_BASE
        rows = _device_rows(lo, hi, dev)
        index = sharded.ReplicatedFlatIndex(rows, "l2", dev)
    else:                        # north-star layout: row shards, top-k allgather, merge kernel
        plan = sharded.ShardPlan(N_BASE, world)
        lo, hi = plan.start(rank), plan.stop(rank)
        rows = _device_rows(lo, hi, dev)
        index = sharded.DistributedFlatIndex(rows, "l2", dev, id_offset=lo)
    del rows
    gq = torch.Generator(device=dev).manual_seed(4242)
    q_dev = torch.randn((NQ, DIM), generator=gq, device=dev, dtype=torch.float32)
    q_host = torch.empty((NQ, DIM), dtype=torch.float32, pin_memory=True)
    q_host.copy_(q_dev)
    q_host_np = q_host.numpy()

    # the reference-facing object for the e2e leg shares this rank's shard (no second copy of the base)
    algo = ExactSearch("exact", DIM, metric="l2")
    algo.index = GpuIndexFlat(DIM, "l2", device=dev)
    algo.index._impl, algo.index.ntotal, algo.index_built = index, N_BASE, True

    shard_bytes = index.memory_bytes()
    flush = shard_bytes < 2 * L2_BYTES
    flush_buf = torch.empty(2 * L2_BYTES // 4, dtype=torch.float32, device=dev) if flush else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device():
        return index.search(q_dev, TOPK)

    # ---- warm-up (also sizes workspaces, creates the NCCL channels); the clock sampler starts here so
    # that the 100 ms nvidia-smi period yields enough samples under load
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # (results are held like in the timed loops: with the previous step's arrays still alive the second
    # call needs a second set of pinned host blocks, and a fresh cudaHostAlloc costs 10-30 ms once)
    for _ in range(max(args.warmup, 3)):
        d_dev, i_dev = step_device()
        d_host, i_host = algo.batch_search(q_host_np, TOPK)
    barrier()

    # ---- timed: device-resident queries
    lib.vdb_flat_timing_enable(1)
    launches0 = lib.vdb_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for e0, e1 in ev:
        if flush:
            flush_buf.fill_(1.0)
        e0.record()
        step_device()
        e1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = lib.vdb_launch_count() - launches0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    import ctypes
    buf = (ctypes.c_float * 512)()
    n_rec = ctypes.c_int(0)
    lib.vdb_flat_timing_read(buf, 512, ctypes.byref(n_rec))
    scan_ms = [buf[i] for i in range(n_rec.value)]
    lib.vdb_flat_timing_enable(0)
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    scan_mean = torch.tensor([sum(scan_ms) / max(len(scan_ms), 1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(scan_mean, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    scan_ms_mean = float(scan_mean.item())

    # the clock sampler covers the device-timed region; it stops here so that nvidia-smi's driver
    # queries cannot stall the synchronous host calls of the end-to-end leg
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed: end to end through ExactSearch.batch_search with host buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d_host, i_host = algo.batch_search(q_host_np, TOPK)
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_s.item()) * 1e3 / args.steps

    # ---- sanity on the result of the last step (not a parity test: tests/ does that)
    assert i_host.shape == (NQ, TOPK) and int(i_host.min()) >= 0 and int(i_host.max()) < N_BASE
    assert bool(np.all(np.diff(d_host, axis=1) >= 0)), "distances are not sorted"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = _peaks()
    # live yardstick on this box: cuBLAS TF32 GEMM 8192^3 (best of 10 after 3 warm-ups, CUDA events)
    cublas_tf32 = None
    try:
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn((8192, 8192), device=dev)
        b = torch.randn((8192, 8192), device=dev)
        for _ in range(3):
            torch.matmul(a, b)
        best = float("inf")
        for _ in range(10):
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(); torch.matmul(a, b); g1.record()
            torch.cuda.synchronize(dev)
            best = min(best, g0.elapsed_time(g1))
        cublas_tf32 = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
        del a, b
    except Exception:      # noqa: BLE001 - yardstick only
        cublas_tf32 = None
    nq_launch = (NQ + world - 1) // world if mode == "queries" else NQ
    flops = 2.0 * nq_launch * (hi - lo) * DIM               # per launch on this rank (2 nq N d)
    pipe_tflops = 3.0 * flops / (scan_ms_mean * 1e-3) / 1e12
    tf32_peak = peaks["bf16_sustained"] / 2.0
    traffic = None
    prof = os.path.join(ROOT, "profiles", "scan_kernel_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:      # noqa: BLE001
            traffic = None

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        nq_sample = 1000
        base_h, q_h = _host_data(nq_sample)
        _cpu_port_qps(base_h, q_h[:50], 1)
        qps_cpu, sec = _cpu_port_qps(base_h, q_h, 2)
        cpu = {"value": qps_cpu, "unit": "queries/s", "cores": _threads(), "kind": "port",
               "sample": f"first {nq_sample} queries against the full 1M x 128 base, best of 2 passes ({sec:.1f} s each); "
                         "oracle FAISS-flat restatement (OpenBLAS sgemm + argpartition)"}

    line = {
        "metric": METRIC, "value": NQ / (ms_per_step * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "tf32x3 (fp32-accurate split) + f64 re-score", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n": N_BASE, "d": DIM, "nq": NQ, "k": TOPK, "metric": "l2",
                   "sharding": ("none" if world == 1 else f"rows/{world}: row shards, NCCL allgather of local top-k, merge kernel"
                                if mode == "rows" else f"queries/{world}: base replicated ({shard_bytes / 1e9:.2f} GB per GPU), "
                                "each rank searches nq/N queries, NCCL allgather of the result blocks"),
                   "l2": ("flushed between steps (256 MB write, outside the per-step events)" if flush else
                          f"operands ({shard_bytes / 1e9:.2f} GB per GPU) exceed the 126 MB L2"),
                   "timing": "per-step CUDA events on the launching stream, summed; max over ranks"},
        "wall_ms_per_step": t_wall * 1e3 / args.steps,
        "e2e": {"value": NQ / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": NQ * DIM * 4, "d2h_bytes_per_step": NQ * TOPK * 12,
                "api": "ExactSearch.batch_search(numpy pinned queries) -> (numpy distances, numpy ids)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "flat_scan_tc_kernel", "achieved": pipe_tflops, "peak": tf32_peak,
                     "unit": "TFLOP/s", "frac": pipe_tflops / tf32_peak, "traffic": traffic,
                     "kernel_ms": scan_ms_mean, "kernel_share_of_step": scan_ms_mean / ms_per_step,
                     "algorithmic_tflops": flops / (scan_ms_mean * 1e-3) / 1e12,
                     "frac_of_burst_peak": pipe_tflops / (peaks["bf16_burst"] / 2.0),
                     "frac_of_nominal_tf32": pipe_tflops / 1125.0,
                     "cublas_tf32_live_tflops": cublas_tf32,
                     "note": f"achieved = 3 * 2*nq*N*d / t of the main scan launch (3xTF32 issues 3 MMAs per product; the "
                             f"seeding pre-pass and the empty redo launch are separate launches inside the step); peak = "
                             f"{peaks['source']} sustained bf16 cuBLAS / 2 (TF32 runs at half the bf16 rate) - a cuBLAS-derived "
                             "proxy the kernel can exceed (frac > 1), so the fraction of the nominal dense TF32 rate "
                             "(1125 TFLOP/s) and a live cuBLAS TF32 8192^3 GEMM are given too; algorithmic_tflops is the "
                             "fp32-equivalent 2*nq*N*d / t"},
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    sys.stdout.flush()
    os.dup2(stdout_fd, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs)")
    ap.add_argument("--shard", choices=["auto", "rows", "queries"], default="auto",
                    help="N > 1: 'rows' = north-star row sharding; 'queries' = replicated base; auto picks by base size")
    args = ap.parse_args()
    if args.steps > 500:
        args.steps = 500
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:      # convenience: re-launch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--impl", args.impl,
               "--shard", args.shard]
        return subprocess.call(cmd + (["--no-cpu-baseline"] if args.no_cpu_baseline else []))
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
