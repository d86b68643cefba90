#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu 2>&1 | tail -40 > gpurun_out/r2c_pytest_conv.txt
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -s 2>&1 | tail -60 > gpurun_out/r2c_pytest_bf16.txt
timeout 900 python -m pytest tests/test_gpu_parity_tc.py tests/test_gpu_modules.py tests/test_mesh_frontend.py tests/test_gpu_ops.py -q -m gpu -s 2>&1 | grep -E "gradient tensors|mask elements|passed|failed|FAILED|Error|error" > gpurun_out/r2c_pytest_parity.txt
cp gpurun_out/parity_report.json gpurun_out/r2c_parity_report.json
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2c_smoke.txt 2>&1
echo finished
