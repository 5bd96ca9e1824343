#!/bin/bash
# round-2 GPU job 6: parity suite after the CRC fold / parser changes, config 2 bench, reduce-side A/B, config 3
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -q --timeout 500 2>&1 | tail -25) > gpurun_out/r2_t6.log
(timeout 600 python bench.py 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_bench6.log
(timeout 300 python tools/merge_profile.py 2>&1 | tail -3 | cut -c1-400) > gpurun_out/r2_merge6_runs.log
(TEZGPU_EMIT_RUNS=0 timeout 300 python tools/merge_profile.py 2>&1 | tail -3 | cut -c1-400) > gpurun_out/r2_merge6_fast4u.log
(timeout 600 python bench.py --config 3 --c3-segments 64 --c3-segment-mb 16 --steps 3 --warmup 1 --c3-cpu-segments 8 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_small6.log
(timeout 1500 python bench.py --config 3 --steps 3 --warmup 1 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_full6.log
tail -8 gpurun_out/r2_t6.log; cat gpurun_out/r2_bench6.log; cat gpurun_out/r2_merge6_runs.log gpurun_out/r2_merge6_fast4u.log; cat gpurun_out/r2_c3_small6.log; cat gpurun_out/r2_c3_full6.log
