#!/bin/bash
# usage: run14.sh N
N=$1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --workload train_step --no-cpu-baseline > gpurun_out/r2n_train_${N}gpu.json 2> gpurun_out/r2n_train_${N}gpu.err
  timeout 600 python bench.py --workload inversion --steps 30 > gpurun_out/r2n_inversion_${N}gpu.json 2> gpurun_out/r2n_inversion_${N}gpu.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload train_step --no-cpu-baseline > gpurun_out/r2n_train_${N}gpu.json 2> gpurun_out/r2n_train_${N}gpu.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload inversion --steps 30 > gpurun_out/r2n_inversion_${N}gpu.json 2> gpurun_out/r2n_inversion_${N}gpu.err
fi
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2n_train_${N}gpu.err | tail -5
python - <<PY
import json
for w in ['train','inversion']:
    try:
        d=json.loads(open('gpurun_out/r2n_%s_${N}gpu.json'%w).read().strip().splitlines()[-1])
        print(w, d['n_gpus'], d['value'], d['unit'], d['ms_per_step'], d.get('collective'), d['clocks'])
    except Exception as e: print(w,'ERR',e)
PY
