#!/bin/bash
# round-2 GPU job 21 (4 GPUs): one config-4 run at N=4
mkdir -p gpurun_out
(timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 2>&1 | tail -n 2 | cut -c1-7000) > gpurun_out/r2_n4.log
cat gpurun_out/r2_n4.log | cut -c1-600
