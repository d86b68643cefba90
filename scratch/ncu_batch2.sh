#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -f -o gpurun_out/prof3_$1 \
      python benchmarks/one_kernel.py $3 4 > gpurun_out/ncu3_$1.log 2>&1
}
cap wgrad128_256 wgrad_tf32_2cta_taps wgrad128_256 2
cap up256_128 conv_igemm_tf32_2cta up256_128 2
ls -la gpurun_out/prof3_*.ncu-rep
