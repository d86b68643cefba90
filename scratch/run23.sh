#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2z_pytest.txt; tail -4 gpurun_out/r2z_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for v in 1 0; do
SR_WEIGHT_PREP_MULTI=$v timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2z_bench_multi$v.json 2> gpurun_out/r2z_bench_multi$v.err
done
python - <<'PY'
import json
for v in '10':
    d=json.loads(open(f'gpurun_out/r2z_bench_multi{v}.json').read())
    a=d['roofline']['all_kernels_ms_per_step']
    print(v, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['gpu_launches'], {k:a[k] for k in a if 'weight' in k})
PY
timeout 600 ncu --set full --clock-control none -k "regex:conv_halo_tf32_2cta_kernel" --launch-skip 60 --launch-count 12 -o gpurun_out/r2_conv_after_fix -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-reference > gpurun_out/r2z_ncu.log 2>&1
tail -1 gpurun_out/r2z_ncu.log
