#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -m gpu -q -x -k "fir_nhwc or fused_pass" 2>&1 | tail -30 > gpurun_out/r2h_pytest_fir.txt
tail -15 gpurun_out/r2h_pytest_fir.txt
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2h_bench_stream.json 2> gpurun_out/r2h_bench_stream.err
SR_FIR_STREAM=0 timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2h_bench_nostream.json 2> gpurun_out/r2h_bench_nostream.err
tail -3 gpurun_out/r2h_bench_stream.err
python - <<'PY'
import json
for f in ['stream','nostream']:
    try:
        d=json.loads(open(f'gpurun_out/r2h_bench_{f}.json').read())
        print(f, d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
        for k,v in d['roofline']['hbm_kernels'].items(): print('   ',k,v)
    except Exception as e: print(f, 'ERR', e)
PY
