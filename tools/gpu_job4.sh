#!/bin/bash
# round-2 GPU job 4 (2 GPUs): config-4 pipeline, pull overlapped with the next sort or not; PCIe ceiling
mkdir -p gpurun_out
(timeout 200 python tools/pcie_probe.py 2>&1 | tail -4) > gpurun_out/r2_pcie.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 6 --warmup 3; }
(TEZ_SHUFFLE_OVERLAP=0 timeout 600 bash -c "$(declare -f run); run 29511" 2>&1 | tail -3) > gpurun_out/r2_n2_seq.log
(TEZ_SHUFFLE_OVERLAP=1 timeout 600 bash -c "$(declare -f run); run 29512" 2>&1 | tail -3) > gpurun_out/r2_n2_overlap.log
cat gpurun_out/r2_pcie.log gpurun_out/r2_n2_seq.log gpurun_out/r2_n2_overlap.log
