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


def test_fetch_segments_verified_copies_and_checks_every_segment():
    """The pull with the IFile CRC32 verified in flight (IFile.Reader.readToMemory semantics, SORT/IFile.java:764-809):
    segments of every size class (empty, shorter than a word, one piece, many 64 KiB pieces) at every alignment."""
    import random
    import zlib
    from oracle import tez_oracle as O
    rng = random.Random(12)
    segs = [O.write_ifile([])[0]]
    for nrec in (1, 2, 7, 100, 3000, 40000):
        recs = sorted((rng.getrandbits(64).to_bytes(8, "big"), rng.randbytes(rng.randint(0, 40))) for _ in range(nrec))
        segs.append(O.write_ifile(recs, rle=False)[0])
    inmem = [s[4:] for s in segs[1:4]]                      # header-less (InMemoryWriter) segments: body + crc
    dev = torch.device("cuda", 0)
    for has_header, group in ((True, segs), (False, inmem)):
        for lead in (0, 1, 4, 7, 12, 15):
            blob = bytearray(b"\xAA" * lead)
            offs = []
            for s_ in group:
                offs.append(len(blob))
                blob += s_                                    # back to back like a file.out range
            src = torch.frombuffer(bytes(blob) + b"\xBB" * 64, dtype=torch.uint8).to(dev)
            dst = torch.full((len(blob) + 64 + 32,), 0xCC, dtype=torch.uint8, device=dev)
            d0 = 16                                           # same residue modulo 16 as the source (both 256-aligned bases)
            table = [(src.data_ptr() + o, dst.data_ptr() + d0 + o, len(s_)) for o, s_ in zip(offs, group)]
            ms = T.fetch_segments_verified(table, has_header=has_header)
            assert ms >= 0
            got = dst.cpu().numpy()
            want = np.full(len(got), 0xCC, dtype=np.uint8)
            want[d0 + lead:d0 + len(blob)] = np.frombuffer(bytes(blob[lead:]), dtype=np.uint8)
            assert np.array_equal(got, want), (has_header, lead)   # segments landed, nothing around them was touched
    # a flipped bit anywhere (body, last partial word, trailer) is a checksum error naming the segment
    big = segs[-1]
    for pos in (4, len(big) // 2, len(big) - 6, len(big) - 1):
        bad = bytearray(big)
        bad[pos] ^= 0x10
        src = torch.frombuffer(bytes(segs[2]) + bytes(bad), dtype=torch.uint8).to(dev)
        dst = torch.zeros(src.numel() + 16, dtype=torch.uint8, device=dev)
        table = [(src.data_ptr(), dst.data_ptr(), len(segs[2])), (src.data_ptr() + len(segs[2]), dst.data_ptr() + len(segs[2]), len(bad))]
        with pytest.raises(IOError, match="checksum mismatch in fetched segment 1"):
            T.fetch_segments_verified(table)


@pytest.mark.parametrize("world,n,P", [(2, 20000, 8), (3, 5000, 7)])
def test_sort_pull_merge_across_processes(world, n, P):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "peer_worker.py"),
           str(n), str(P), "3"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == world
