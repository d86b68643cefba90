#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2m_pytest.txt
tail -12 gpurun_out/r2m_pytest.txt
timeout 600 python benchmarks/train_step.py --profile --iters 16 > gpurun_out/r2m_train_profile.txt 2> gpurun_out/r2m_train_profile.err
cut -c1-300 gpurun_out/r2m_train_profile.txt
