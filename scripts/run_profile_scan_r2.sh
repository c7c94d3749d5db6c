#!/bin/bash
# Round-2 full capture of the dominant kernel (main scan launch of the C2 bench) -> profiles/scan_kernel_r2.md + traffic json
TAG=${1:-r2z}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/prof_scan_plain_$TAG.json 2> gpurun_out/prof_scan_plain_$TAG.err &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:flat_scan_tc_kernel<.*128, .*, \(int\)0>" --launch-skip 8 --launch-count 1 -o gpurun_out/prof_scan_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_scan_$TAG.log 2>&1; echo "scan capture rc=$?"
ls -la gpurun_out/prof_scan_$TAG.ncu-rep
