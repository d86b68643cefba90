#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2i_bench_stream.json 2> gpurun_out/r2i_bench_stream.err
python - <<'PY'
import json
for f in ['r2i_bench_stream']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read())
        print(f, d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
        for k,v in d['roofline']['hbm_kernels'].items(): print('   ',k,v)
        for k,v in d['roofline']['tensor_shapes'].items(): print('   ',k,v)
    except Exception as e: print(f, 'ERR', e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:fir_nhwc_stream_kernel" --launch-skip 16 --launch-count 8 -o gpurun_out/r2_fir_stream -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-reference > gpurun_out/r2i_ncu.log 2>&1
tail -3 gpurun_out/r2i_ncu.log
