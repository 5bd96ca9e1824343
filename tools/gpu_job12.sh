#!/bin/bash
# round-2 GPU job 12: parity suite with the evidence-ranked guess, config 5 (N=1) and config 3 again
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60) > gpurun_out/r2_t12.log
(timeout 900 python bench.py --config 5 --steps 4 --warmup 3 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_c5_n1_12.log
(timeout 600 python bench.py --config 3 --c3-segments 64 --c3-segment-mb 16 --steps 3 --warmup 1 --c3-cpu-segments 8 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_small12.log
(timeout 1500 python bench.py --config 3 --steps 3 --warmup 1 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_full12.log
(timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 700 --csv --log-file gpurun_out/r2_launches_c5_12.csv python bench.py --config 5 --steps 1 --warmup 1 > /dev/null 2>&1)
tail -6 gpurun_out/r2_t12.log; cat gpurun_out/r2_c5_n1_12.log; cat gpurun_out/r2_c3_small12.log; cat gpurun_out/r2_c3_full12.log; (timeout 600 python tools/e2e_probe.py 2>&1 | tail -40) > gpurun_out/r2_e2e_probe.log; cat gpurun_out/r2_e2e_probe.log
