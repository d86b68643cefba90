#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_modules.py tests/test_gpu_parity_tc.py tests/test_gpu_bf16.py -m gpu -q -x 2>&1 | tail -15
for v in 1 0; do
  SR_BLUR_CONV=$v timeout 600 python bench.py --workload train_step --no-cpu-baseline > gpurun_out/r2r_train_blurconv$v.json 2> gpurun_out/r2r_train_blurconv$v.err
  tail -2 gpurun_out/r2r_train_blurconv$v.err
done
python - <<'PY'
import json
for v in '10':
    d=json.loads(open(f'gpurun_out/r2r_train_blurconv{v}.json').read().strip().splitlines()[-1])
    print(v, d['value'], d['ms_per_step'], d['gpu_launches'], d['losses'])
    print({k:v for k,v in d['roofline']['all_kernels_ms_per_step'].items() if v>0.3})
PY
