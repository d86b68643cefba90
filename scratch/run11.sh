#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2k_pytest.txt
tail -8 gpurun_out/r2k_pytest.txt
timeout 600 python benchmarks/gwm_bench.py > gpurun_out/r2k_gwm.json 2> gpurun_out/r2k_gwm.err; tail -3 gpurun_out/r2k_gwm.err; cat gpurun_out/r2k_gwm.json | cut -c1-1500
timeout 600 python bench.py --workload train_step --no-cpu-baseline > gpurun_out/r2k_train_1gpu.json 2> gpurun_out/r2k_train_1gpu.err; tail -3 gpurun_out/r2k_train_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_train_1gpu.json').read())
print(d['value'], d['ms_per_step'], d['clocks'], d['gpu_launches'], d.get('losses'))
print(d['roofline']['all_kernels_ms_per_step'])
PY
