#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2e_pytest_all.txt
timeout 300 python bench.py --workload rasterize --steps 20 > gpurun_out/r2e_raster.json 2> gpurun_out/r2e_raster.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2e_bench_sep.json 2> gpurun_out/r2e_bench_sep.err
SR_FIR_SEP=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2e_bench_nosep.json 2> gpurun_out/r2e_bench_nosep.err
SR_PROLOGUE_SPEC=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2e_bench_spec2.json 2> gpurun_out/r2e_bench_spec2.err
timeout 600 python benchmarks/train_step.py --iters 8 --profile --no-e2e > gpurun_out/r2e_train_profile.json 2> gpurun_out/r2e_train_profile.txt
timeout 600 python bench.py --workload inversion --steps 30 > gpurun_out/r2e_inversion.json 2> gpurun_out/r2e_inversion.err
echo finished
