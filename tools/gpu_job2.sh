#!/bin/bash
# round-2 GPU job 2: full parity suite, TMA-emit subset, reduce-side profile, full default bench line
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 2>&1 | tail -30) > gpurun_out/r2_t2.log
(TEZGPU_EMIT_TMA=1 timeout 600 python -m pytest tests/test_sorter_gpu.py -m gpu -q --timeout 200 2>&1 | tail -5) > gpurun_out/r2_t2_tma.log
(timeout 300 python tools/merge_profile.py 2>&1 | tail -20) > gpurun_out/r2_merge_profile2.log
(timeout 900 python bench.py 2>&1 | tail -3) > gpurun_out/r2_bench_n1.log
tail -6 gpurun_out/r2_t2.log; tail -3 gpurun_out/r2_t2_tma.log; cat gpurun_out/r2_merge_profile2.log; cat gpurun_out/r2_bench_n1.log
