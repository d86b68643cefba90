#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -40 > gpurun_out/r2b_pytest_all.txt
timeout 900 python -m pytest tests/test_gpu_parity_tc.py -q -m gpu -s 2>&1 | grep -E "gradient tensors|passed|failed|FAILED|Error" > gpurun_out/r2b_pytest_parity.txt
cp gpurun_out/parity_report.json gpurun_out/r2b_parity_report.json
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2b_smoke.txt 2>&1
timeout 300 python bench.py --workload rasterize --steps 20 > gpurun_out/r2b_raster.json 2> gpurun_out/r2b_raster.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 300 python benchmarks/gwm_bench.py > gpurun_out/r2b_gwm.json 2> gpurun_out/r2b_gwm.err
timeout 600 python bench.py --workload train_step > gpurun_out/r2b_train_1gpu.json 2> gpurun_out/r2b_train_1gpu.err
echo finished
