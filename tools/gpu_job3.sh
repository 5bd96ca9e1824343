#!/bin/bash
# round-2 GPU job 3: config-3 shape parity tests, config 3 bench (small then full), sorter bench sanity
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_merger_gpu.py tests/test_runtime_library_gpu.py tests/test_sorter_gpu.py -m gpu -q --timeout 300 2>&1 | tail -15) > gpurun_out/r2_t3.log
(timeout 600 python bench.py --config 3 --c3-segments 64 --c3-segment-mb 16 --steps 3 --warmup 1 --c3-cpu-segments 8 2>&1 | tail -n 30 | cut -c1-6000) > gpurun_out/r2_c3_small.log
(timeout 1500 python bench.py --config 3 --steps 3 --warmup 1 2>&1 | tail -n 30 | cut -c1-6000) > gpurun_out/r2_c3_full.log
(timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-g1-pipeline 2>&1 | tail -2) > gpurun_out/r2_bench_n1_b.log
tail -8 gpurun_out/r2_t3.log; cat gpurun_out/r2_c3_small.log; cat gpurun_out/r2_c3_full.log; cat gpurun_out/r2_bench_n1_b.log
