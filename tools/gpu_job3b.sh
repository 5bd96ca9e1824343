#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python tools/diag_c3.py 2>&1) > gpurun_out/r2_diag_c3.log
(timeout 300 python tools/merge_profile.py 2>&1 | tail -8) > gpurun_out/r2_merge_profile3.log
(timeout 300 python -m pytest tests/test_merger_gpu.py tests/test_peer_fetch_gpu.py -m gpu -q --timeout 300 2>&1 | tail -5) > gpurun_out/r2_t3b.log
cat gpurun_out/r2_diag_c3.log gpurun_out/r2_merge_profile3.log gpurun_out/r2_t3b.log
