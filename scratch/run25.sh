#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_conv.py -m gpu -q -x -k "double_backward or conv_tc" 2>&1 | tail -12
timeout 600 python bench.py --workload train_step --no-cpu-baseline > gpurun_out/r3c_train.json 2> gpurun_out/r3c_train.err; tail -2 gpurun_out/r3c_train.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3c_train.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['phase_ms'], d['losses'])
PY
