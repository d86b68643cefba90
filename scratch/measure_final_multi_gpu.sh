#!/bin/bash
N=$1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
if [ "$N" = "1" ]; then L=""; else L="-m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
timeout 900 python $L bench.py --gpus $N --workload train_step --no-cpu-baseline > gpurun_out/r2x_train_${N}gpu.json 2> gpurun_out/r2x_train_${N}gpu.err
timeout 900 python $L bench.py --gpus $N --workload inversion --steps 30 > gpurun_out/r2x_inversion_${N}gpu.json 2> gpurun_out/r2x_inversion_${N}gpu.err
timeout 900 python $L bench.py --gpus $N --steps 30 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2x_generator_${N}gpu.json 2> gpurun_out/r2x_generator_${N}gpu.err
python - <<PY
import json
for w in ['train','inversion','generator']:
    try:
        d=json.loads(open('gpurun_out/r2x_%s_${N}gpu.json'%w).read().strip().splitlines()[-1])
        print(w, d['n_gpus'], d['value'], d['unit'], d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'), d['clocks']['sm_mhz'])
    except Exception as e: print(w,'ERR',e)
PY
