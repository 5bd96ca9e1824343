#!/bin/bash
# round-2 GPU job 7: parity suite, config 2 bench (ILP off in the capped kernels), reduce-side A/B, config 3 with the agreeing-guess parser
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -q --timeout 500 2>&1 | tail -25) > gpurun_out/r2_t7.log
(timeout 600 python bench.py 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_bench7.log
(timeout 300 python tools/merge_profile.py 2>&1 | tail -2 | cut -c1-400) > gpurun_out/r2_merge7_runs.log
(TEZGPU_EMIT_RUNS=0 timeout 300 python tools/merge_profile.py 2>&1 | tail -2 | cut -c1-400) > gpurun_out/r2_merge7_fast4u.log
(timeout 600 python bench.py --config 3 --c3-segments 64 --c3-segment-mb 16 --steps 3 --warmup 1 --c3-cpu-segments 8 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_small7.log
(timeout 1500 python bench.py --config 3 --steps 3 --warmup 1 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_full7.log
(timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_c3_small.csv python bench.py --config 3 --c3-segments 64 --c3-segment-mb 16 --steps 1 --warmup 0 --c3-cpu-segments 0 > /dev/null 2>&1)
tail -8 gpurun_out/r2_t7.log; cat gpurun_out/r2_bench7.log; cat gpurun_out/r2_merge7_runs.log gpurun_out/r2_merge7_fast4u.log; cat gpurun_out/r2_c3_small7.log; cat gpurun_out/r2_c3_full7.log
