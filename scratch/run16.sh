#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2p_pytest.txt
tail -8 gpurun_out/r2p_pytest.txt
timeout 600 python benchmarks/kernel_bench.py --out gpurun_out/r2p_kernel_bench.jsonl > /dev/null 2> gpurun_out/r2p_kernel_bench.err
tail -2 gpurun_out/r2p_kernel_bench.err
grep -E "upfirdn2d_blur|fused_bias|rasterize" gpurun_out/r2p_kernel_bench.jsonl | cut -c1-260
