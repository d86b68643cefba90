#!/bin/bash
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload train_step --no-cpu-baseline > gpurun_out/r2f_train_2gpu.json 2> gpurun_out/r2f_train_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload inversion --steps 30 > gpurun_out/r2f_inversion_2gpu.json 2> gpurun_out/r2f_inversion_2gpu.err
timeout 600 python bench.py --workload train_step --no-cpu-baseline > gpurun_out/r2f_train_1gpu.json 2> gpurun_out/r2f_train_1gpu.err
timeout 600 python bench.py --workload inversion --steps 30 > gpurun_out/r2f_inversion_1gpu.json 2> gpurun_out/r2f_inversion_1gpu.err
tail -3 gpurun_out/r2f_*.err
echo finished
