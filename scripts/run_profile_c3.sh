#!/bin/bash
# ncu --set full captures of the HBM-bound kernels on config C3 (one launch each), after a plain run.
# usage: scripts/run_profile_c3.sh <tag> [with_hamming]
TAG=${1:-r1c}
set -x
python scripts/bench_c3.py > gpurun_out/c3_$TAG.jsonl 2> gpurun_out/c3_$TAG.err; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ivf_scan_kernel --launch-skip 37 --launch-count 1 -o gpurun_out/prof_ivf_scan_$TAG -f python scripts/bench_c3.py > gpurun_out/ncu_ivf.log 2>&1; echo "ivf rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rerank_topk_kernel --launch-skip 13 --launch-count 1 -o gpurun_out/prof_rerank_$TAG -f python scripts/bench_c3.py > gpurun_out/ncu_rerank.log 2>&1; echo "rerank rc=$?"
if [ -n "$2" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hamming_emit_kernel --launch-skip 9 --launch-count 1 -o gpurun_out/prof_hamming_$TAG -f python scripts/bench_c3.py > gpurun_out/ncu_hamming.log 2>&1; echo "hamming rc=$?"
fi
ls -la gpurun_out/prof_*_$TAG.ncu-rep
