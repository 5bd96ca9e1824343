#!/bin/bash
# round-2 GPU job 4 (2 GPUs): PCIe ceiling, config-4 pipeline with the pull overlapped with the next sort or not, config 5 at N=2
mkdir -p gpurun_out
(timeout 200 python tools/pcie_probe.py 2>&1 | tail -4) > gpurun_out/r2_pcie.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 6 --warmup 3 $2 $3; }
(TEZ_SHUFFLE_OVERLAP=0 timeout 600 bash -c "$(declare -f run); run 29511" 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_n2_seq.log
(TEZ_SHUFFLE_OVERLAP=1 timeout 600 bash -c "$(declare -f run); run 29512" 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_n2_overlap.log
(timeout 600 bash -c "$(declare -f run); run 29513 --config 5" 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_n2_c5.log
(timeout 600 bash -c "$(declare -f run); run 29514 --impl reference" 2>&1 | tail -n 2 | cut -c1-3000) > gpurun_out/r2_n2_ref.log
cat gpurun_out/r2_pcie.log; for f in r2_n2_seq r2_n2_overlap r2_n2_c5 r2_n2_ref; do echo "== $f"; cat gpurun_out/$f.log | cut -c1-2500; done
