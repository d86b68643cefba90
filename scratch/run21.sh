#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_bf16.py tests/test_gpu_parity_tc.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
timeout 600 python bench.py --workload rasterize --steps 20 > gpurun_out/r2w_raster.json 2> gpurun_out/r2w_raster.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read())
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['peak'], d['roofline']['frac'])
print({k:v for k,v in d['roofline']['all_kernels_ms_per_step'].items() if v>0.3})
for k,v in list(d['roofline']['tensor_shapes'].items())[:10]: print('   ',k,v)
r=json.loads(open('gpurun_out/r2w_raster.json').read())
print(r['value'], r['ms_per_step'], r['roofline'])
PY
