"""GPU parity of the unordered writer mode (UnorderedPartitionedKVWriter behind
UnorderedPartitionedKVOutput): partition only, no key order.  The oracle restates the single-buffer case of
RL/common/writers/UnorderedPartitionedKVWriter.java:459-472,688-703,1058-1144 (newest record first per partition, IFile
without run-length encoding, all-zero index entries for partitions without records); the device output is compared
byte for byte, index included."""
import random

import numpy as np
import pytest

from oracle import tez_oracle as O
import tez_b200 as T

pytestmark = pytest.mark.gpu


def _pack(records):
    kv = bytearray()
    ko, vo, vl = [], [], []
    for k, v in records:
        ko.append(len(kv))
        kv += k
        vo.append(len(kv))
        kv += v
        vl.append(len(v))
    return (np.frombuffer(bytes(kv), dtype=np.uint8) if kv else np.zeros(0, np.uint8), np.array(ko, np.uint32),
            np.array(vo, np.uint32), np.array(vl, np.uint32))


@pytest.mark.parametrize("n", [0, 1, 2, 300, 5000, 200000])
@pytest.mark.parametrize("P", [1, 7, 64])
def test_unordered_variable_length_records_bit_exact(n, P):
    rng = random.Random(n * 131 + P)
    recs = [(rng.randbytes(rng.randint(0, 12)), rng.randbytes(rng.randint(0, 40))) for _ in range(n)]
    kv, ko, vo, vl = _pack(recs)
    for given in (False, True):
        part = np.array([rng.randrange(P) for _ in range(n)], np.int32) if given else None
        conf = O.sorter_conf(P, cmp_kind=O.CMP_BYTES, partitioner=O.PART_GIVEN if given else O.PART_HASH)
        exp = O.unordered_write(conf, kv, ko.astype(np.uint64), vo - ko, vl, part)
        with T.GpuSorter(P, comparator=T.CMP_BYTES, partitioner=T.PART_GIVEN if given else T.PART_HASH, unordered=True) as s:
            half = n // 2
            for a, b in ((0, half), (half, n)):          # two collects: arrival order spans batches
                if b > a:
                    lo, hi = int(ko[a]), int(vo[b - 1] + vl[b - 1])
                    s.collect(kv[lo:hi], ko[a:b] - lo, vo[a:b] - lo, vl[a:b], None if part is None else part[a:b])
            out, index_bytes, index, st = s.flush_to_memory()
        assert bytes(out) == exp["file_out"]
        assert index_bytes == exp["index_out"]
        assert np.array_equal(index, exp["index"])
        assert st["output_records"] == n and not st["rle_used"]
        assert st["output_bytes_with_overhead"] == exp["counters"]["OUTPUT_BYTES_WITH_OVERHEAD"]


@pytest.mark.parametrize("n,P", [(1000, 4), (1000003, 64), (300000, 1024)])
def test_unordered_fixed_width_records_bit_exact(n, P):
    """The same through the fixed-width fast path (16 B key / 64 B value): the TMA-free aligned emit kernels."""
    kv = O.gen_c2(0, n, seed=5)
    ko = np.arange(n, dtype=np.uint64) * 80
    exp = O.unordered_write(O.sorter_conf(P), kv, ko, np.full(n, 16, np.uint32), np.full(n, 64, np.uint32))
    with T.GpuSorter(P, fixed=(16, 64), unordered=True) as s:
        s.collect_fixed(kv)
        out, index_bytes, index, st = s.flush_to_memory()
    assert np.array_equal(np.frombuffer(bytes(out), dtype=np.uint8), np.frombuffer(exp["file_out"], dtype=np.uint8))
    assert index_bytes == exp["index_out"]
    # multiset per partition == the ordered sorter's (the two writers only differ in the order inside a segment)
    srt = O.pipelined_sort_fixed(O.sorter_conf(P, rle_policy=0), kv, 16, 64)
    assert np.array_equal(srt["index"][:, 1][srt["index"][:, 1] > 6], exp["index"][:, 1][exp["index"][:, 1] > 0])
