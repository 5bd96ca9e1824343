#!/bin/bash
# one GPU-side validation pass of the current tree: full GPU suite, the opt-in emit variants, A/B timings
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
TEZGPU_EMIT_ROUND_FILL=1 python -m pytest tests/test_sorter_gpu.py tests/test_merger_gpu.py tests/test_runtime_library_gpu.py -x -q 2>&1 | tail -1
TEZGPU_EMIT_PIPE_UNALIGNED=1 python -m pytest tests/test_merger_gpu.py tests/test_peer_fetch_gpu.py tests/test_runtime_library_gpu.py tests/test_sorter_gpu.py -x -q 2>&1 | tail -1
python tools/ab_emit.py default
TEZGPU_EMIT_ROUND_FILL=1 python tools/ab_emit.py default
TEZGPU_EMIT_PIPE_UNALIGNED=1 python tools/merge_profile.py 2>&1 | tail -1
