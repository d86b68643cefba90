#!/bin/bash
mkdir -p gpurun_out
for spec in default 2; do
  if [ "$spec" = "2" ]; then export SR_PROLOGUE_SPEC=2; else unset SR_PROLOGUE_SPEC; fi
  timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2q_bench_spec_$spec.json 2> gpurun_out/r2q_bench_spec_$spec.err
done
unset SR_PROLOGUE_SPEC
python - <<'PY'
import json
for f in ['default','2']:
    d=json.loads(open(f'gpurun_out/r2q_bench_spec_{f}.json').read())
    print(f, d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['hbm_kernels']['sr_styled_bwd_prologue3_f32'])
d=json.loads(open('gpurun_out/r2q_bench_spec_default.json').read())
for k,v in list(d['roofline']['tensor_shapes'].items())[:14]: print('   ',k,v)
PY
timeout 300 python -m pytest tests/test_gpu_modules.py -m gpu -q -x -k "fused_pass or chain or styled" 2>&1 | tail -3
