#!/bin/bash
# round-2 GPU job 18: parity suite + config 3 / 5 / 1 after the repeat-flag pass
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40) > gpurun_out/r2_t18.log
(timeout 900 python bench.py --config 3 --steps 3 --warmup 1 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_full18.log
(timeout 600 python bench.py --config 5 --steps 4 --warmup 3 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_c5_n1_18.log
(timeout 600 python bench.py --config 1 2>&1 | tail -n 3 | cut -c1-5000) > gpurun_out/r2_c1_18.log
tail -5 gpurun_out/r2_t18.log; cat gpurun_out/r2_c3_full18.log gpurun_out/r2_c5_n1_18.log gpurun_out/r2_c1_18.log | cut -c1-1200
