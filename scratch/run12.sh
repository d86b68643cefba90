#!/bin/bash
mkdir -p gpurun_out
timeout 600 python benchmarks/train_step.py --profile --iters 16 > gpurun_out/r2l_train_profile.txt 2> gpurun_out/r2l_train_profile.err
tail -3 gpurun_out/r2l_train_profile.err
grep -v "^--" gpurun_out/r2l_train_profile.txt | cut -c1-100,190-260 | head -60
