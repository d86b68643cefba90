#!/bin/bash
# round-2 GPU call 1: parity tests (new first), smoke, bench A/B of the prologue specialisation, train_step on 1 GPU
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity_tc.py -q -m gpu 2>&1 | tail -60 > gpurun_out/r2_pytest_parity.txt
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity_tc.py 2>&1 | tail -40 > gpurun_out/r2_pytest_rest.txt
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2_smoke.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
SR_PROLOGUE_SPEC=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2_bench_spec0.json 2> gpurun_out/r2_bench_spec0.err
SR_PROLOGUE_SPEC=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2_bench_spec2.json 2> gpurun_out/r2_bench_spec2.err
timeout 600 python bench.py --workload train_step --no-cpu-baseline > gpurun_out/r2_train_1gpu.json 2> gpurun_out/r2_train_1gpu.err
timeout 300 python bench.py --workload rasterize --steps 20 > gpurun_out/r2_raster_a.json 2> gpurun_out/r2_raster_a.err
echo finished
