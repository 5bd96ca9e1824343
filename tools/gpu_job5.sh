#!/bin/bash
# round-2 GPU job 5: full parity suite, config 3 (small, full), ncu of the reduce-side emit kernels
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 2>&1 | tail -25) > gpurun_out/r2_t5.log
(timeout 600 python bench.py --config 3 --c3-segments 64 --c3-segment-mb 16 --steps 3 --warmup 1 --c3-cpu-segments 8 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_small.log
(timeout 1500 python bench.py --config 3 --steps 3 --warmup 1 2>&1 | tail -n 12 | cut -c1-5000) > gpurun_out/r2_c3_full.log
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_emit_runs -c 1 -o gpurun_out/r2_prof_emit_runs python tools/merge_profile.py 20000000 2 1024 2>&1 | tail -3) > gpurun_out/r2_ncu_runs.log
(TEZGPU_EMIT_RUNS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_emit_fast4u -c 1 -o gpurun_out/r2_prof_emit_fast4u python tools/merge_profile.py 20000000 2 1024 2>&1 | tail -3) > gpurun_out/r2_ncu_fast4u.log
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2_launches_merge.csv python tools/merge_profile.py 100000000 2 1024 > /dev/null 2>&1)
for g in 0 64 128; do (TEZGPU_L2_FETCH=$g timeout 300 python tools/merge_profile.py 2>&1 | tail -2 | cut -c1-400) > gpurun_out/r2_merge_l2fetch_$g.log; done
for g in 0 64 128; do (TEZGPU_L2_FETCH=$g TEZGPU_EMIT_RUNS=0 timeout 300 python tools/merge_profile.py 2>&1 | tail -1 | cut -c1-400) > gpurun_out/r2_merge_l2fetch_fast4u_$g.log; done
tail -6 gpurun_out/r2_t5.log; tail -n 3 gpurun_out/r2_merge_l2fetch_*.log; cat gpurun_out/r2_c3_small.log; cat gpurun_out/r2_c3_full.log; cat gpurun_out/r2_ncu_runs.log gpurun_out/r2_ncu_fast4u.log; ls -la gpurun_out/*.ncu-rep | tail -3
