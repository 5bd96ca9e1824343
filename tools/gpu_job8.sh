#!/bin/bash
# round-2 GPU job 8: parity suite, config 3 with the guess/evaluate/chase parser, config 1 and config 5 lines
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -80) > gpurun_out/r2_t9.log
(timeout 600 python bench.py --config 3 --c3-segments 64 --c3-segment-mb 16 --steps 3 --warmup 1 --c3-cpu-segments 8 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_small9.log
(timeout 1500 python bench.py --config 3 --steps 3 --warmup 1 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_full9.log
(timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3_small.csv python bench.py --config 3 --c3-segments 64 --c3-segment-mb 16 --steps 1 --warmup 1 --c3-cpu-segments 2 > /dev/null 2>&1)
(timeout 900 python bench.py --config 1 2>&1 | tail -n 6 | cut -c1-5000) > gpurun_out/r2_c1_9.log
(timeout 900 python bench.py --config 5 --steps 4 --warmup 3 2>&1 | tail -n 6 | cut -c1-6000) > gpurun_out/r2_c5_n1_9.log
(timeout 300 python tools/merge_profile.py 2>&1 | tail -2 | cut -c1-400) > gpurun_out/r2_merge9.log
tail -8 gpurun_out/r2_t9.log; cat gpurun_out/r2_c3_small9.log; cat gpurun_out/r2_c3_full9.log; cat gpurun_out/r2_c1_9.log gpurun_out/r2_c5_n1_9.log gpurun_out/r2_merge9.log
