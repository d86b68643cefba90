#!/bin/bash
cd "$(dirname "$0")/.."
for shape in "256 128 128 up" "512 256 64 up" "256 128 128 gather" "512 256 64 gather"; do
  for dbg in 0 1 3; do
    SR_CONV_DEBUG=$dbg timeout 120 python benchmarks/conv_probe.py $shape 2>&1 | tail -1
  done
  SR_CONV_HALO=0 timeout 120 python benchmarks/conv_probe.py $shape 2>&1 | tail -1
done
