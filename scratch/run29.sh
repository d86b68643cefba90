#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_parity_tc.py tests/test_gpu_conv.py tests/test_gpu_bf16.py -m gpu -q -x -k "discriminator or plain_conv or conv_tc or bf16" 2>&1 | tail -4
for v in 1 0; do
SR_RES_COMBINE=$v timeout 600 python bench.py --workload train_step --no-cpu-baseline > gpurun_out/r3g_train$v.json 2> gpurun_out/r3g_train$v.err
python -c "
import json; d=json.loads(open('gpurun_out/r3g_train$v.json').read().strip().splitlines()[-1]); print($v, d['value'], d['ms_per_step'], d['phase_ms'], d['losses'])"
done
