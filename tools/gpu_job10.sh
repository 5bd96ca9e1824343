#!/bin/bash
# round-2 GPU job 10: parity suite, config 3 with the tail-first guess, launch lists (config 3 small, config 5, merge with DRAM bytes)
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60) > gpurun_out/r2_t10.log
(timeout 600 python bench.py --config 3 --c3-segments 64 --c3-segment-mb 16 --steps 3 --warmup 1 --c3-cpu-segments 8 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_small10.log
(timeout 1500 python bench.py --config 3 --steps 3 --warmup 1 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_full10.log
(timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r2_launches_c3_small10.csv python bench.py --config 3 --c3-segments 64 --c3-segment-mb 16 --steps 1 --warmup 1 --c3-cpu-segments 2 > /dev/null 2>&1)
(timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 700 --csv --log-file gpurun_out/r2_launches_c5.csv python bench.py --config 5 --steps 1 --warmup 1 > gpurun_out/r2_c5_ncu.log 2>&1)
(timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -s 30 -c 120 --csv --log-file gpurun_out/r2_launches_merge10.csv python tools/merge_profile.py 100000000 2 1024 > /dev/null 2>&1)
tail -6 gpurun_out/r2_t10.log; cat gpurun_out/r2_c3_small10.log; cat gpurun_out/r2_c3_full10.log; tail -3 gpurun_out/r2_c5_ncu.log | cut -c1-600
