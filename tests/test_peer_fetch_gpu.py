"""NVLink pull shuffle: the fetch kernel (any alignment) and the multi-process sort -> pull -> merge path.
The multi-process test shares cuda:0 between the ranks (CUDA IPC maps a buffer of the same device just as well), so
it runs on a one-GPU box; bench.py --gpus N covers the one-rank-per-GPU placement."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import tez_b200 as T

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fetch_ranges_every_alignment_and_size():
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(3)
    src = torch.randint(0, 256, (3 << 20,), dtype=torch.uint8, generator=g).to(dev)
    dst = torch.zeros(4 << 20, dtype=torch.uint8, device=dev)
    ranges, expect = [], []
    cur = 0
    cases = [(0, 0, 0), (1, 1, 1), (5, 5, 10), (15, 15, 17), (3, 3, 31), (0, 0, 16), (7, 7, 4096 + 9), (2, 2, 300000),
             (9, 9, 131072 * 3 + 5), (4, 11, 1000), (1, 0, 70000), (0, 0, 1 << 20)]
    soff = 0
    for smis, dmis, ln in cases:
        s = (soff + 15) // 16 * 16 + smis
        d = (cur + 15) // 16 * 16 + dmis
        ranges.append((src.data_ptr() + s, dst.data_ptr() + d, ln))
        expect.append((s, d, ln))
        soff, cur = s + ln, d + ln + 3   # 3 guard bytes after every destination range
    ms = T.fetch_ranges(ranges)
    assert ms >= 0
    h_src, h_dst = src.cpu().numpy(), dst.cpu().numpy()
    want = np.zeros_like(h_dst)
    for s, d, ln in expect:
        want[d:d + ln] = h_src[s:s + ln]
    assert np.array_equal(h_dst, want)   # ranges landed, guard bytes and gaps untouched


@pytest.mark.parametrize("world,n,P", [(2, 20000, 8), (3, 5000, 7)])
def test_sort_pull_merge_across_processes(world, n, P):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "peer_worker.py"),
           str(n), str(P), "3"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == world
