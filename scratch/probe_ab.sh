#!/bin/bash
# A/B inside one box: barrier polling mode (4) and direct epilogue stores (8) of the halo kernel, two repetitions
cd "$(dirname "$0")/.."
for rep in 1 2; do
for shape in "128 128 256" "256 256 128"; do
  for dbg in 0 4 8 12; do
    SR_CONV_DEBUG=$dbg timeout 120 python benchmarks/conv_probe.py $shape plain 2>&1 | tail -1
  done
done
done
