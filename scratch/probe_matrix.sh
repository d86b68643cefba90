#!/bin/bash
# conv bottleneck probe: im2col vs halo (one box per class / per dx), with the epilogue and/or the MMAs disabled
cd "$(dirname "$0")/.."
for shape in "128 128 256" "512 512 64"; do
for cfg in "0 1" "1 0" "1 1"; do
  set -- $cfg
  for dbg in 0 1 2 3; do
    SR_CONV_HALO=$1 SR_HALO_SPLITX=$2 SR_CONV_DEBUG=$dbg timeout 120 python benchmarks/conv_probe.py $shape plain 2>&1 | tail -1
  done
done
done
