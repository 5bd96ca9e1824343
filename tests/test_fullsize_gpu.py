"""GPU, BASELINE full size (config 2: 1e8 records, 16 B key / 64 B value, 64 partitions): size-independent properties
of the device output, checked on the device with torch (the oracle takes minutes at this size):
  sortedness inside every partition, partition == HashPartitioner(key), multiset of records preserved,
  every segment header / EOF / CRC32 (zlib on the host), index arithmetic of TezIndexRecord."""
import zlib

import numpy as np
import pytest
import torch

import tez_b200 as T
from tez_b200 import synth

pytestmark = pytest.mark.gpu

N, P = 100_000_000, 64


def _fold(x):
    """order-independent 64-bit fingerprint of the rows of an int64 matrix"""
    h = torch.zeros(x.shape[0], dtype=torch.int64, device=x.device)
    for j in range(x.shape[1]):
        h = synth.splitmix64(h ^ x[:, j])
    return int(h.sum().item())


def test_config2_full_size_properties():
    dev = torch.device("cuda", 0)
    d_kv = synth.gen_c2(0, N, seed=2, device=dev)
    cap = N * 82 + 10 * P + 4096
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    with T.GpuSorter(P, fixed=(16, 64)) as s:
        out_len, index, st = s.sort_device_fixed(d_kv.data_ptr(), N, d_out.data_ptr(), cap)
    assert st["output_records"] == N and not st["rle_used"]
    assert out_len == N * 82 + 10 * int((index[:, 1] > 0).sum())
    # fingerprint of the input records (as 10 big-endian words each)
    words_in = d_kv.view(N, 80)
    fp_in = 0
    for a in range(0, N, 1 << 24):
        blk = words_in[a:a + (1 << 24)].view(-1, 10, 8).to(torch.int64)
        w = torch.zeros(blk.shape[0], 10, dtype=torch.int64, device=dev)
        for b in range(8):
            w = (w << 8) | blk[:, :, b]
        fp_in = (fp_in + _fold(w)) & ((1 << 64) - 1)
    del words_in
    fp_out, total, off = 0, 0, 0
    for p in range(P):
        start, raw, part = (int(x) for x in index[p])
        assert start == off and part == raw + 4
        off += part
        cnt = (raw - 6) // 82
        assert raw == 6 + 82 * cnt
        total += cnt
        seg = d_out[start:start + part]
        host_seg = seg.cpu().numpy()
        assert bytes(host_seg[:4]) == b"TIF\x00" and bytes(host_seg[-6:-4]) == b"\xff\xff"
        assert int.from_bytes(bytes(host_seg[-4:]), "big") == zlib.crc32(host_seg[4:-4])
        recs = seg[4:4 + 82 * cnt].view(cnt, 82)
        assert bool((recs[:, 0] == 16).all()) and bool((recs[:, 1] == 64).all())
        keys = recs[:, 2:18]
        assert bool((synth.hash_partition(keys, P) == p).all())
        k = keys.to(torch.int64)
        hi = torch.zeros(cnt, dtype=torch.int64, device=dev)
        lo = torch.zeros(cnt, dtype=torch.int64, device=dev)
        for b in range(8):
            hi = (hi << 8) | k[:, b]
            lo = (lo << 8) | k[:, 8 + b]
        # unsigned order: flip the sign bit
        hi, lo = hi ^ (-1 << 63), lo ^ (-1 << 63)
        ok = (hi[:-1] < hi[1:]) | ((hi[:-1] == hi[1:]) & (lo[:-1] <= lo[1:]))
        assert bool(ok.all()), "partition %d not sorted" % p
        blk = recs[:, 2:].reshape(cnt, 10, 8).to(torch.int64)
        w = torch.zeros(cnt, 10, dtype=torch.int64, device=dev)
        for b in range(8):
            w = (w << 8) | blk[:, :, b]
        fp_out = (fp_out + _fold(w)) & ((1 << 64) - 1)
        del recs, keys, k, hi, lo, blk, w
    assert total == N and off == out_len
    assert fp_in & ((1 << 64) - 1) == fp_out, "records were lost or altered"
