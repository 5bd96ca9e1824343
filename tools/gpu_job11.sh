#!/bin/bash
# round-2 GPU job 11: parity suite (wire format tests, tie-fix change), config 2 bench with the direction-locked e2e leg, config 3
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60) > gpurun_out/r2_t11.log
(timeout 900 python bench.py 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_bench11.log
(timeout 600 python bench.py --config 3 --c3-segments 64 --c3-segment-mb 16 --steps 3 --warmup 1 --c3-cpu-segments 8 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_small11.log
(timeout 1500 python bench.py --config 3 --steps 3 --warmup 1 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_full11.log
tail -6 gpurun_out/r2_t11.log; cat gpurun_out/r2_bench11.log; cat gpurun_out/r2_c3_small11.log; cat gpurun_out/r2_c3_full11.log
