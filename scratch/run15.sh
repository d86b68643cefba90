#!/bin/bash
N=$1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --workload train_step --no-cpu-baseline > gpurun_out/r2o_train_${N}gpu.json 2> gpurun_out/r2o_train_${N}gpu.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload train_step --no-cpu-baseline > gpurun_out/r2o_train_${N}gpu.json 2> gpurun_out/r2o_train_${N}gpu.err
fi
grep -v "^\*\|OMP_NUM\|^$\|Grad strides\|grad.sizes\|bucket_view\|run_backward" gpurun_out/r2o_train_${N}gpu.err | tail -12
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2o_train_${N}gpu.json').read().strip().splitlines()[-1])
    print(d['n_gpus'], d['value'], d['unit'], d['ms_per_step'], d['e2e'], d.get('collective'), d['clocks'], d['gpu_launches'], d['losses'])
except Exception as e: print('ERR',e)
PY
