#!/bin/bash
# round-2 GPU job 15 (8 GPUs): one config-4 run at N=8 (1.25e8 records per GPU = 1e9 records), sequential exchange
mkdir -p gpurun_out
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 2>&1 | tail -n 3 | cut -c1-7000) > gpurun_out/r2_n8b.log
cat gpurun_out/r2_n8b.log | cut -c1-3000
