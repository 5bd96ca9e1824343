#!/bin/bash
# round-2 GPU job 19: the N=1 line and the reference arm at the round's final build
mkdir -p gpurun_out
(timeout 600 python bench.py 2>&1 | tail -n 3 | cut -c1-7000) > gpurun_out/r2_final_bench_b.log
(timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -n 2 | cut -c1-4000) > gpurun_out/r2_final_ref_b.log
cat gpurun_out/r2_final_bench_b.log gpurun_out/r2_final_ref_b.log | cut -c1-1500
