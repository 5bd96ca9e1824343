"""BASELINE config 1 driver (tools/owc_bench.cc): OrderedWordCount, two ordered edges, known answer.
CPU: the job through the oracle arm.  GPU: the same job through the plugin mirror over the CUDA library."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exe():
    sys.path.insert(0, ROOT)
    import __graft_entry__
    from oracle import tez_oracle
    tez_oracle.build()
    return __graft_entry__.build_tools()


def test_ordered_word_count_known_answer_cpu_arm(tmp_path):
    out = subprocess.check_output([_exe(), "cpu", "3", "3", "4", str(tmp_path)], text=True)
    r = json.loads(out.strip().splitlines()[-1])
    assert r["answer_checked"] and r["records"] == (3 << 20) // 7 and r["reducers"] == 4


@pytest.mark.gpu
def test_ordered_word_count_known_answer_gpu_arm(tmp_path):
    out = subprocess.check_output([_exe(), "gpu", "8", "3", "4", str(tmp_path)], text=True)
    r = json.loads(out.strip().splitlines()[-1])
    assert r["answer_checked"] and r["records"] == (8 << 20) // 7
