#!/bin/bash
# C2 strong-scaling spot check on one multi-GPU box: bench.py at the given GPU counts (auto layout), plus rows at the largest.
# usage: scripts/run_scaling_c2.sh "4 8"
for N in $1; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_c2b_n$N.json 2> gpurun_out/scale_c2b_n$N.err
  python -c "import json;d=json.load(open('gpurun_out/scale_c2b_n$N.json'));print('N=$N auto', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'kernel', round(d['roofline']['kernel_ms'],3))"
  LAST=$N
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $LAST --master-addr 127.0.0.1 --master-port 2966 bench.py --gpus $LAST --steps 20 --warmup 5 --shard rows > gpurun_out/scale_c2b_rows_n$LAST.json 2> gpurun_out/scale_c2b_rows_n$LAST.err
python -c "import json;d=json.load(open('gpurun_out/scale_c2b_rows_n$LAST.json'));print('N=$LAST rows', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'kernel', round(d['roofline']['kernel_ms'],3))"
