#!/bin/bash
# 8-GPU evidence run: multi-GPU parity tests, bench.py at N=8 (rows headline + replicated), C5 100M x 128 and C4 8.8M x 768 at N=8
N=${1:-8}
python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/scale8_pytest_multigpu.log 2>&1; tail -3 gpurun_out/scale8_pytest_multigpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale8_c2_n$N.json 2> gpurun_out/scale8_c2_n$N.err
$TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --shard rows --exchange allgather --no-parity > gpurun_out/scale8_c2_allgather_n$N.json 2>> gpurun_out/scale8_c2_n$N.err
$TR --master-port 29513 scripts/scale_100m.py --steps 3 > gpurun_out/scale8_100m_n$N.json 2> gpurun_out/scale8_100m_n$N.err
$TR --master-port 29514 scripts/scale_100m.py --steps 3 --rows 8800000 --dim 768 --metric ip > gpurun_out/scale8_c4_n$N.json 2> gpurun_out/scale8_c4_n$N.err
for f in c2 c2_allgather; do python - <<PY
import json
d=json.load(open('gpurun_out/scale8_${f}_n$N.json'))
print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'parity', d['parity'] and d['parity']['ok'], 'repl', d.get('replicated') and (round(d['replicated']['value']), round(d['replicated']['e2e']['value'])))
PY
done
tail -1 gpurun_out/scale8_100m_n$N.json | cut -c1-600; tail -1 gpurun_out/scale8_c4_n$N.json | cut -c1-600
