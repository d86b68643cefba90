#!/bin/bash
# ncu --set full captures of the main kernels (one launch each), reports land in gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {  # name regex one_kernel-arg skip
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -f -o gpurun_out/prof2_$1 \
      python benchmarks/one_kernel.py $3 4 > gpurun_out/ncu2_$1.log 2>&1
}
cap conv512_64 conv_halo conv512_64 2
cap conv128_256 conv_halo conv128_256 2
cap wgrad512_64 wgrad_tf32_2cta wgrad512_64 2
cap wgrad128_256 "wgrad_tf32_kernel" wgrad128_256 2
cap up256_128 conv_igemm_tf32_2cta up256_128 2
cap blur_nchw upfirdn2d_tile blur_nchw 2
cap blur_nhwc upfirdn2d_nhwc blur_nhwc 2
cap prologue styled_bwd_prologue prologue 2
cap raster_tri raster_tri raster 2
cap raster_resolve raster_resolve raster 2
cap raster_bwd raster_backward raster 2
cap bias_act bias_act_vec4 bias_act 2
ls -la gpurun_out/prof2_*.ncu-rep
