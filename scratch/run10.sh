#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_bf16.py tests/test_gpu_parity_tc.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2j_pytest.txt
tail -8 gpurun_out/r2j_pytest.txt
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2j_bench_stream.json 2> gpurun_out/r2j_bench_stream.err
python - <<'PY'
import json
for f in ['r2j_bench_stream']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read())
        print(f, d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
        for k,v in d['roofline']['hbm_kernels'].items(): print('   ',k,v)
    except Exception as e: print(f, 'ERR', e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:fir_nhwc_stream_kernel" --launch-skip 16 --launch-count 8 -o gpurun_out/r2_fir_stream2 -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-reference > gpurun_out/r2j_ncu.log 2>&1
tail -2 gpurun_out/r2j_ncu.log
