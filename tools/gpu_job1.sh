#!/bin/bash
# round-2 GPU job 1: micro-benchmark of the TMA gather, parity suites (register-staged emit, then TMA emit), A/B bench
mkdir -p gpurun_out
(timeout 200 tools/bench_gather 2>&1) > gpurun_out/r2_gather1.log
(TEZGPU_EMIT_TMA=0 timeout 1200 python -m pytest tests -m gpu -q --timeout 400 2>&1 | tail -40) > gpurun_out/r2_t1_notma.log
(timeout 600 python -m pytest tests/test_sorter_gpu.py tests/test_fullsize_gpu.py tests/test_runtime_library_gpu.py -m gpu -q --timeout 200 2>&1 | tail -40) > gpurun_out/r2_t1_tma.log
(TEZGPU_EMIT_TMA=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e 2>&1 | tail -3) > gpurun_out/r2_bench_notma.log
(timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e 2>&1 | tail -3) > gpurun_out/r2_bench_tma.log
(timeout 300 python tools/merge_profile.py 2>&1 | tail -20) > gpurun_out/r2_merge_profile.log
tail -4 gpurun_out/r2_t1_notma.log; tail -4 gpurun_out/r2_t1_tma.log; cat gpurun_out/r2_gather1.log; cat gpurun_out/r2_bench_notma.log gpurun_out/r2_bench_tma.log; cat gpurun_out/r2_merge_profile.log
