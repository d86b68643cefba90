#!/bin/bash
mkdir -p gpurun_out
# launch list of one eager step window (per-launch durations, cold cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 900 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2v_launch.log 2>&1
tail -1 gpurun_out/r2v_launch.log | cut -c1-200
# full capture of the dominant conv launch (128 -> 128 @ 256^2) and the prologue
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_halo_tf32_2cta_kernel|styled_bwd_prologue_kernel" --launch-skip 60 --launch-count 24 -o gpurun_out/r2_conv_prologue -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-reference > gpurun_out/r2v_ncu.log 2>&1
tail -2 gpurun_out/r2v_ncu.log
