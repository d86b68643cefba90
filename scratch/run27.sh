#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_tc.py tests/test_mesh_frontend.py -m gpu -q -x -k "raster or mesh" 2>&1 | tail -4
for v in 1 0; do
SR_RASTER_COMPACT=$v timeout 300 python bench.py --workload rasterize --steps 20 --no-cpu-baseline > gpurun_out/r3e_raster$v.json 2> gpurun_out/r3e_raster$v.err
python -c "
import json; d=json.loads(open('gpurun_out/r3e_raster$v.json').read()); print($v, d['value'], d['ms_per_step'], d['roofline']['forward_ms'], d['roofline']['backward_ms'], d['roofline']['frac'])"
done
