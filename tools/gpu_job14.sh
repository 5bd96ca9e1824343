#!/bin/bash
# round-2 GPU job 14: e2e timeline probe; launch list of one full config-3 step
mkdir -p gpurun_out
(timeout 600 python tools/e2e_probe.py 2>&1 | tail -40) > gpurun_out/r2_e2e_probe.log
(timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/r2_launches_c3_full.csv python bench.py --config 3 --steps 1 --warmup 1 --c3-cpu-segments 2 > /dev/null 2>&1)
cat gpurun_out/r2_e2e_probe.log
