#!/bin/bash
# Round-1 closing run on one B200: parity suite, smoke, bench line, launch list, C5 at one GPU.
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_final.log
tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1d_ref.json 2> gpurun_out/bench_r1d_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_d.log 2>&1; echo "ncu launches rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_r1d.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['clocks'])"
cat gpurun_out/bench_r1d_ref.json | cut -c1-300
