#!/bin/bash
# single-CTA im2col kernel (no cluster protocol): same probe
cd "$(dirname "$0")/.."
for shape in "128 128 256" "512 512 64" "256 256 128"; do
  for dbg in 0 1 2 3; do
    SR_CONV_HALO=0 SR_CONV_2CTA=0 SR_CONV_DEBUG=$dbg timeout 120 python benchmarks/conv_probe.py $shape plain 2>&1 | tail -1
  done
done
