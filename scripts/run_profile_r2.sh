#!/bin/bash
# Round-2 ncu evidence (one GPU): launch list of the headline bench, full captures of the Hamming scan (C = 6 400 and 800),
# the rerank at C = 800 and the IVF scan at nprobe = 1.  Each ncu run follows a plain run of the same command (exit 0).
# usage: scripts/run_profile_r2.sh <tag>
TAG=${1:-r2}
NCU="ncu --clock-control none"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/prof_bench_plain_$TAG.json 2> gpurun_out/prof_bench_plain_$TAG.err &&
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launch list rc=$?"
python scripts/bench_c3.py > gpurun_out/c3_$TAG.jsonl 2> gpurun_out/c3_$TAG.err; echo "plain c3 rc=$?"
HAM='regex:flat_scan_tc_kernel<.*, \(int\)1>'
timeout 600 $NCU --set full --import-source on --kernel-name-base demangled -k "$HAM" --launch-skip 9 --launch-count 1 -o gpurun_out/prof_hamming_tc_c6400_$TAG -f python scripts/bench_c3.py > gpurun_out/ncu_ham1.log 2>&1; echo "hamming 6400 rc=$?"
timeout 600 $NCU --set full --import-source on --kernel-name-base demangled -k "$HAM" --launch-skip 1 --launch-count 1 -o gpurun_out/prof_hamming_tc_c800_$TAG -f python scripts/bench_c3.py > gpurun_out/ncu_ham2.log 2>&1; echo "hamming 800 rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:rerank_topk_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/prof_rerank_c800_$TAG -f python scripts/bench_c3.py > gpurun_out/ncu_rr.log 2>&1; echo "rerank 800 rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:ivf_scan_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/prof_ivf_np1_$TAG -f python scripts/bench_c3.py > gpurun_out/ncu_ivf.log 2>&1; echo "ivf nprobe 1 rc=$?"
ls -la gpurun_out/*_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
