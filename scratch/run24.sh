#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_parity_tc.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --workload train_step --no-cpu-baseline > gpurun_out/r3a_train.json 2> gpurun_out/r3a_train.err; tail -2 gpurun_out/r3a_train.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3a_train.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['phase_ms'])
print({k:v for k,v in d['roofline']['all_kernels_ms_per_step'].items() if v>0.3})
PY
timeout 300 python benchmarks/gwm_bench.py 2>/dev/null | tail -1
