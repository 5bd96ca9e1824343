"""Seeded, counter-based synthetic workloads generated directly with torch so that the benchmark
inputs can be produced in HBM (or pinned host memory) without shipping data.  Bit-identical to the oracle's
generators (tests/test_synth.py checks that on CPU)."""
import torch

_M1 = -0x61C8864680B583EB  # 0x9E3779B97F4A7C15 as int64
_M2 = -0x40A7B892E31B1A47  # 0xBF58476D1CE4E5B9
_M3 = -0x6B2FB644ECCEEE15  # 0x94D049BB133111EB


def _lsr(x, s):
    return (x >> s) & ((1 << (64 - s)) - 1)


def splitmix64(x):
    """x: int64 tensor (two's complement view of uint64). Wrap-around arithmetic == uint64 arithmetic."""
    x = x + _M1
    x = (x ^ _lsr(x, 30)) * _M2
    x = (x ^ _lsr(x, 27)) * _M3
    return x ^ _lsr(x, 31)


def _be_bytes(words):
    """int64 tensor [..., W] -> uint8 tensor [..., W*8] big-endian."""
    shifts = torch.arange(56, -8, -8, device=words.device, dtype=torch.int64)
    b = (words.unsqueeze(-1) >> shifts) & 0xFF
    return b.to(torch.uint8).reshape(*words.shape[:-1], -1)


def gen_c2(first_index, n, seed=2, device="cpu", out=None, chunk=1 << 22):
    """C2 records: 80 bytes = 16 B key (2 big-endian splitmix64 words) + 64 B value (8 words); word w of record i is
    splitmix64((seed << 56) ^ (16 * i + w)).  Returns a uint8 tensor of n*80 bytes."""
    if out is None:
        out = torch.empty(n * 80, dtype=torch.uint8, device=device)
    dev = out.device
    w = torch.arange(10, device=dev, dtype=torch.int64)
    sbits = (seed << 56)
    if sbits >= 1 << 63:
        sbits -= 1 << 64
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        i = torch.arange(first_index + s, first_index + s + m, device=dev, dtype=torch.int64)
        words = splitmix64((i.unsqueeze(1) * 16 + w) ^ sbits)
        out[s * 80:(s + m) * 80] = _be_bytes(words).reshape(-1)
    return out


def hash_bytes(keys):
    """WritableComparator.hashBytes over the rows of a uint8 tensor [n, L] -> int32-valued int64 tensor."""
    h = torch.ones(keys.shape[0], dtype=torch.int64, device=keys.device)
    sb = keys.to(torch.int64)
    sb = torch.where(sb > 127, sb - 256, sb)
    for j in range(keys.shape[1]):
        h = (h * 31 + sb[:, j]) & 0xFFFFFFFF
    return h


def hash_partition(keys, num_partitions):
    return (hash_bytes(keys) & 0x7FFFFFFF) % num_partitions
