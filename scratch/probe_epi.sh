#!/bin/bash
cd "$(dirname "$0")/.."
for shape in "128 128 256 plain" "256 256 128 plain" "512 512 64 plain" "256 128 128 up" "512 256 64 up"; do
  for e in 4 8; do
    SR_CONV_EPI_WARPS=$e timeout 120 python benchmarks/conv_probe.py $shape 2>&1 | tail -1 | sed "s/$/ EPI=$e/"
  done
done
