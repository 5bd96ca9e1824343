#!/bin/bash
# round-2 GPU job 20: ncu --set full of the general emit kernel (config 3's dominant kernel) on the 1 GB variant
mkdir -p gpurun_out
(timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:^k_emit$' -s 1 -c 1 -f -o gpurun_out/r2_prof_emit_var python bench.py --config 3 --c3-segments 64 --c3-segment-mb 16 --steps 1 --warmup 1 --c3-cpu-segments 2 2>&1 | tail -4) > gpurun_out/r2_ncu_emit_var.log
cat gpurun_out/r2_ncu_emit_var.log | cut -c1-400; ls -la gpurun_out/r2_prof_emit_var.ncu-rep
