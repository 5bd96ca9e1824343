#!/bin/bash
# last GPU call of the round: smoke, full parity suite, the N=1 bench line and the reference arm
mkdir -p gpurun_out
(timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3) > gpurun_out/r2_final_smoke.log
(timeout 1700 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40) > gpurun_out/r2_final_tests.log
(timeout 900 python bench.py 2>&1 | tail -n 3 | cut -c1-7000) > gpurun_out/r2_final_bench.log
(timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -n 2 | cut -c1-4000) > gpurun_out/r2_final_ref.log
cat gpurun_out/r2_final_smoke.log; tail -5 gpurun_out/r2_final_tests.log; cat gpurun_out/r2_final_bench.log; cat gpurun_out/r2_final_ref.log
