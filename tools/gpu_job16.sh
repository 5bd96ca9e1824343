#!/bin/bash
# round-2 GPU job 16: parity suite; e2e timeline (small results stored by kernel); config 2 bench; config 3 (warp metadata emit); config 5 (hinted guess)
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60) > gpurun_out/r2_t16.log
(timeout 600 python tools/e2e_probe.py 2>&1 | tail -50) > gpurun_out/r2_e2e_probe16.log
(timeout 900 python bench.py 2>&1 | tail -n 3 | cut -c1-7000) > gpurun_out/r2_bench16.log
(timeout 1500 python bench.py --config 3 --steps 3 --warmup 1 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_full16.log
(timeout 900 python bench.py --config 5 --steps 4 --warmup 3 2>&1 | tail -n 3 | cut -c1-6000) > gpurun_out/r2_c5_n1_16.log
tail -6 gpurun_out/r2_t16.log; cat gpurun_out/r2_e2e_probe16.log; cat gpurun_out/r2_bench16.log; cat gpurun_out/r2_c3_full16.log; cat gpurun_out/r2_c5_n1_16.log
