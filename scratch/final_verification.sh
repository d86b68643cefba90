#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r3d_pytest.txt; tail -3 gpurun_out/r3d_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r3d_bench_default.json 2> gpurun_out/r3d_bench_default.err
timeout 600 python bench.py --workload train_step --no-cpu-baseline > gpurun_out/r3d_train.json 2> gpurun_out/r3d_train.err
timeout 600 python bench.py --workload inversion --steps 30 > gpurun_out/r3d_inversion.json 2> gpurun_out/r3d_inversion.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3d_reference_arm.json 2> gpurun_out/r3d_reference_arm.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3d_bench_default.json').read())
print('generator', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['gpu_launches'], d['roofline']['frac'], d.get('gpu_reference',{}).get('value'), d.get('cpu_baseline',{}).get('value'))
print({k:(v['ms_per_step'],v['frac']) for k,v in d['roofline']['hbm_kernels'].items()})
t=json.loads(open('gpurun_out/r3d_train.json').read().strip().splitlines()[-1])
print('train', t['value'], t['ms_per_step'], t['phase_ms'])
print({k:(v['ms_per_step'],v['frac']) for k,v in t['roofline']['hbm_kernels'].items()})
i=json.loads(open('gpurun_out/r3d_inversion.json').read().strip().splitlines()[-1])
print('inversion', i['value'], i['ms_per_step'])
r=json.loads(open('gpurun_out/r3d_reference_arm.json').read().strip().splitlines()[-1])
print('reference arm', r.get('value'), r.get('impl'))
PY
