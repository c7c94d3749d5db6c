#!/bin/bash
# Scaling sweep on one box: bench.py (C2, strong scaling) and the 100M x 128 sweep at N = 1, 2, 4, 8.
# usage: scripts/run_scaling.sh <max_gpus> ; outputs gpurun_out/scale_*.json
MAXN=${1:-8}
SKIP100M=${2:-0}
for N in 1 2 4 8; do
  [ $N -gt $MAXN ] && break
  if [ $N -eq 1 ]; then
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_c2_n$N.json 2> gpurun_out/scale_c2_n$N.err
    [ $SKIP100M -eq 0 ] && python scripts/scale_100m.py --steps 3 > gpurun_out/scale_100m_n$N.json 2> gpurun_out/scale_100m_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_c2_n$N.json 2> gpurun_out/scale_c2_n$N.err
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 297$N bench.py --gpus $N --steps 20 --warmup 5 --shard rows > gpurun_out/scale_c2rows_n$N.json 2> gpurun_out/scale_c2rows_n$N.err
    [ $SKIP100M -eq 0 ] && python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 296$N scripts/scale_100m.py --steps 3 > gpurun_out/scale_100m_n$N.json 2> gpurun_out/scale_100m_n$N.err
    echo "N=$N c2 rows: $(python -c "import json;d=json.load(open('gpurun_out/scale_c2rows_n$N.json'));print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))" 2>&1 | tail -1)"
  fi
  echo "N=$N c2: $(python -c "import json;d=json.load(open('gpurun_out/scale_c2_n$N.json'));print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))" 2>&1 | tail -1)"
  [ $SKIP100M -eq 0 ] && echo "N=$N 100m: $(tail -1 gpurun_out/scale_100m_n$N.json | cut -c1-300)"
done
