#!/bin/bash
python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py -m gpu -q -x -k "upfirdn2d or fir_nhwc or fused_pass" 2>&1 | tail -4
python scratch/planes_sweep.py 2>&1 | tail -9
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
for k,v in d['roofline']['hbm_kernels'].items(): print('   ',k,v)
PY
