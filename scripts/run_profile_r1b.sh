set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_r1b.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_b.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_scan_tc --launch-skip 4 --launch-count 1 -o gpurun_out/prof_scan_r1b -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_b.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
