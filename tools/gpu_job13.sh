#!/bin/bash
# round-2 GPU job 13 (2 GPUs): overlap with the pull grid capped at one CTA per SM vs sequential; config 5 at N=2
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 6 --warmup 3 $2 $3; }
(TEZ_SHUFFLE_OVERLAP=1 timeout 600 bash -c "$(declare -f run); run 29512" 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_n2_overlap148.log
(TEZ_SHUFFLE_OVERLAP=1 TEZGPU_FETCH_CTAS=74 timeout 600 bash -c "$(declare -f run); run 29515" 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_n2_overlap74.log
(TEZ_SHUFFLE_OVERLAP=0 TEZ_SHUFFLE_VERIFY_IN_FETCH=0 timeout 600 bash -c "$(declare -f run); run 29516" 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_n2_seq_noverify.log
(timeout 600 bash -c "$(declare -f run); run 29513 --config 5" 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_n2_c5_13.log
for f in r2_n2_overlap148 r2_n2_overlap74 r2_n2_seq_noverify r2_n2_c5_13; do echo "== $f"; cat gpurun_out/$f.log | cut -c1-1500; done
