#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s 2>&1 | grep -E "gradient tensors|mask elements|passed|failed|FAILED|Error|error" > gpurun_out/r2d_pytest_all.txt
cp gpurun_out/parity_report.json gpurun_out/r2d_parity_report.json
timeout 600 python bench.py --workload train_step --precision bf16 > gpurun_out/r2d_train_bf16.json 2> gpurun_out/r2d_train_bf16.err
timeout 600 python bench.py --workload train_step --precision tf32 --no-cpu-baseline > gpurun_out/r2d_train_tf32.json 2> gpurun_out/r2d_train_tf32.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster -c 8 -o gpurun_out/r2_raster python bench.py --workload rasterize --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_ncu_raster.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:upfirdn2d_nhwc_kernel|styled_bwd_prologue_kernel" --launch-skip 75 --launch-count 25 -o gpurun_out/r2_hbm_passes python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-reference > gpurun_out/r2d_ncu_hbm.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo finished
