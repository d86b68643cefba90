#!/bin/bash
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3f_bench_2gpu.json 2> gpurun_out/r3f_bench_2gpu.err; echo rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r3f_ref_2gpu.json 2> gpurun_out/r3f_ref_2gpu.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3f_bench_2gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['clocks']['sm_mhz'])
r=json.loads(open('gpurun_out/r3f_ref_2gpu.json').read().strip().splitlines()[-1])
print(r.get('impl'), r.get('value'), r.get('n_gpus'))
PY
grep -v "Warning\|warn\|^\*\|OMP" gpurun_out/r3f_bench_2gpu.err | tail -3 | cut -c1-200
