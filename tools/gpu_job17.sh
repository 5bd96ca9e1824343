#!/bin/bash
# round-2 GPU job 17: config 3 and config 5 at the final build (metadata pass in groups of eight lanes)
mkdir -p gpurun_out
(timeout 1500 python bench.py --config 3 --steps 3 --warmup 1 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_full17.log
(timeout 600 python bench.py --config 3 --c3-segments 64 --c3-segment-mb 16 --steps 3 --warmup 1 --c3-cpu-segments 8 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_small17.log
(timeout 900 python bench.py --config 5 --steps 4 --warmup 3 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_c5_n1_17.log
cat gpurun_out/r2_c3_full17.log gpurun_out/r2_c3_small17.log gpurun_out/r2_c5_n1_17.log | cut -c1-1500
