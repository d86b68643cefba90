#!/bin/bash
cd "$(dirname "$0")/.."
for shape in "256 128 128 up" "512 256 64 up" "512 512 32 up" "256 128 128 gather" "512 256 64 gather" "512 512 8 plain" "512 512 4 plain"; do
  timeout 120 python benchmarks/conv_probe.py $shape 2>&1 | tail -1
done
