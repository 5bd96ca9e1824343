"""GPU parity tests: libtezgpu (CUDA, through the C ABI) vs the CPU oracle, bit-exact on the same seeded inputs.
Mirrors the reference's TestPipelinedSorter / TestDefaultSorter cases (basic, empty, RLE, counters, file perms)."""
import os
import random
import stat
import zlib

import numpy as np
import pytest

from oracle import tez_oracle as O
import tez_b200 as T

pytestmark = pytest.mark.gpu


def _oracle_conf(P, cmp_kind, part, rle, send_empty=True, legacy=False):
    return O.sorter_conf(P, cmp_kind=cmp_kind, partitioner=part, send_empty=send_empty, rle_policy=rle, legacy=legacy)


def _pack(records):
    """records: list of (key bytes, value bytes) -> kv buffer + offsets"""
    kv = bytearray()
    ko, vo, vl = [], [], []
    for k, v in records:
        ko.append(len(kv))
        kv += k
        vo.append(len(kv))
        kv += v
        vl.append(len(v))
    return (np.frombuffer(bytes(kv), dtype=np.uint8) if kv else np.zeros(0, np.uint8),
            np.array(ko, np.uint32), np.array(vo, np.uint32), np.array(vl, np.uint32))


def _check_against_oracle(records, P, cmp_kind, rle=-1, partition=None, send_empty=True, legacy=False, batches=1):
    kv, ko, vo, vl = _pack(records)
    part_mode = O.PART_GIVEN if partition is not None else O.PART_HASH
    exp = O.pipelined_sort(_oracle_conf(P, cmp_kind, part_mode, rle, send_empty, legacy), kv, ko.astype(np.uint64),
                           vo - ko, vl, partition)
    with T.GpuSorter(P, comparator=cmp_kind, partitioner=part_mode, rle_policy=rle, send_empty=send_empty,
                     legacy=legacy) as s:
        n = len(records)
        step = max(1, (n + batches - 1) // batches)
        for a in range(0, n, step):
            b = min(n, a + step)
            lo = int(ko[a])
            hi = int(vo[b - 1] + vl[b - 1])
            s.collect(kv[lo:hi], ko[a:b] - lo, vo[a:b] - lo, vl[a:b],
                      None if partition is None else np.asarray(partition[a:b], np.int32))
        out, index_bytes, index, st = s.flush_to_memory()
    assert bytes(out) == exp["file_out"], "file.out differs from the oracle"
    assert index_bytes == exp["index_out"]
    assert np.array_equal(index, exp["index"])
    assert bool(st["rle_used"]) == exp["rle_used"]
    c = exp["counters"]
    assert st["output_records"] == c["OUTPUT_RECORDS"]
    assert st["output_bytes"] == c["OUTPUT_BYTES"]
    assert st["output_bytes_with_overhead"] == c["OUTPUT_BYTES_WITH_OVERHEAD"]
    assert st["output_bytes_physical"] == c["OUTPUT_BYTES_PHYSICAL"]
    assert st["spilled_records"] == c["SPILLED_RECORDS"]
    assert st["num_spills"] == 1 and st["kernel_launches"] > 0
    return exp, st


@pytest.mark.parametrize("n", [0, 1, 2, 255, 256, 257, 5000, 100000, 1000003])
@pytest.mark.parametrize("P", [1, 64])
def test_c2_fixed_width_bit_exact(n, P):
    """BASELINE config 2 shape: 16 B key / 64 B value, HashPartitioner, TezBytesComparator."""
    kv = O.gen_c2(0, n, seed=2)
    exp = O.pipelined_sort_fixed(O.sorter_conf(P), kv, 16, 64)
    with T.GpuSorter(P, fixed=(16, 64)) as s:
        s.collect_fixed(kv)
        out, index_bytes, index, st = s.flush_to_memory()
    assert bytes(out) == exp["file_out"]
    assert index_bytes == exp["index_out"]
    assert st["output_records"] == n and not st["rle_used"]
    assert st["output_bytes"] == n * 80


@pytest.mark.parametrize("n", [1_000_000, 10_000_000])
def test_c4_shape_1024_partitions_bit_exact(n):
    """BASELINE config 4's per-GPU map side: 16 B key / 64 B value into 1024 partitions, compared byte for byte with the
    oracle at sizes it still finishes in seconds (file.out, file.out.index, every partition's index triple)."""
    import torch
    kv = O.gen_c2(0, n, seed=11)
    exp = O.pipelined_sort_fixed(O.sorter_conf(1024), kv, 16, 64)
    with T.GpuSorter(1024, fixed=(16, 64)) as s:
        d_kv = torch.from_numpy(kv).cuda()
        cap = n * 82 + 10 * 1024 + 4096
        d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        out_len, index, st = s.sort_device_fixed(d_kv.data_ptr(), n, d_out.data_ptr(), cap)
        out = d_out[:out_len].cpu().numpy()
    assert out_len == len(exp["file_out"])
    assert np.array_equal(out, np.frombuffer(exp["file_out"], dtype=np.uint8)), "file.out differs from the oracle"
    assert np.array_equal(index, exp["index"])
    assert st["output_records"] == n and not st["rle_used"]


def test_look_back_kernels_make_progress_next_to_other_streams():
    """The chained-scan kernels (radix_sort.cuh, k_part_bounds) wait on the tile before them and identify tiles by
    blockIdx: forward progress relies on CTAs of ONE grid being dispatched in index order, also when another stream's
    kernels occupy SMs (the multi-GPU step runs the peer pull next to the sort).  Two competing streams keep every SM
    busy with long kernels while sorts of several sizes run; every result must still equal the oracle's."""
    import threading
    import torch
    n = 3_000_000
    kv = O.gen_c2(0, n, seed=21)
    exp = O.pipelined_sort_fixed(O.sorter_conf(64), kv, 16, 64)
    stop = threading.Event()

    def hog():
        torch.cuda.set_device(0)
        st = torch.cuda.Stream()
        a = torch.randn(4096, 4096, device="cuda")
        junk = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
        with torch.cuda.stream(st):
            while not stop.is_set():
                for _ in range(4):
                    a = torch.tanh(a @ a * 1e-3)      # full-grid GEMMs + elementwise kernels
                    junk.add_(1)                      # bandwidth-bound grid of > 1e5 CTAs
                st.synchronize()

    threads = [threading.Thread(target=hog) for _ in range(2)]
    for t in threads:
        t.start()
    try:
        with T.GpuSorter(64, fixed=(16, 64)) as s:
            d_kv = torch.from_numpy(kv).cuda()
            cap = n * 82 + 64 * 16 + 4096
            d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
            for it in range(12):
                out_len, index, st = s.sort_device_fixed(d_kv.data_ptr(), n, d_out.data_ptr(), cap)
                assert out_len == len(exp["file_out"])
                if it % 4 == 0:
                    assert np.array_equal(d_out[:out_len].cpu().numpy(), np.frombuffer(exp["file_out"], dtype=np.uint8))
                assert np.array_equal(index, exp["index"])
    finally:
        stop.set()
        for t in threads:
            t.join()


def test_c2_fixed_multiple_collects_and_files(tmp_path):
    n, P = 30000, 7
    kv = O.gen_c2(100, n, seed=9)
    exp = O.pipelined_sort_fixed(O.sorter_conf(P), kv, 16, 64)
    out_path, idx_path = str(tmp_path / "file.out"), str(tmp_path / "file.out.index")
    with T.GpuSorter(P, fixed=(16, 64)) as s:
        for a in range(0, n, 7001):
            s.collect_fixed(kv[a * 80:min(n, a + 7001) * 80])
        index, st = s.flush(out_path, idx_path)
    assert open(out_path, "rb").read() == exp["file_out"]
    assert open(idx_path, "rb").read() == exp["index_out"]
    # RLT/.../TestPipelinedSorter.java:889-898 -- spill files are 0640
    assert stat.S_IMODE(os.stat(out_path).st_mode) == 0o640
    assert stat.S_IMODE(os.stat(idx_path).st_mode) == 0o640
    assert np.array_equal(index, exp["index"])


def test_fixed_width_with_given_partitions_and_odd_sizes():
    rng = np.random.default_rng(5)
    n, P, kl, vl = 20000, 13, 10, 7   # unaligned stride: generic (non-FAST16) staging and unaligned gathers
    kv = rng.integers(0, 256, size=n * (kl + vl), dtype=np.uint8)
    part = rng.integers(0, P, size=n, dtype=np.int32)
    conf = O.sorter_conf(P, partitioner=O.PART_GIVEN)
    ko = np.arange(n, dtype=np.uint64) * (kl + vl)
    exp = O.pipelined_sort(conf, kv, ko, np.full(n, kl, np.uint32), np.full(n, vl, np.uint32), part)
    with T.GpuSorter(P, fixed=(kl, vl), partitioner=T.PART_GIVEN) as s:
        s.collect_fixed(kv, part)
        out, index_bytes, _, _ = s.flush_to_memory()
    assert bytes(out) == exp["file_out"] and index_bytes == exp["index_out"]


def _text_records(words, value_fn):
    return [(O.text(w), value_fn(w)) for w in words]


def test_wordcount_like_text_keys_rle_on():
    """BASELINE config 1 shape: (Text word, IntWritable 1), few distinct words => heavy duplication => RLE on."""
    rng = random.Random(1)
    vocab = ["a_%d" % i for i in range(1, 11)] + ["w%05d" % i for i in range(50)]
    words = [rng.choice(vocab) for _ in range(20000)]
    exp, st = _check_against_oracle(_text_records(words, lambda w: O.int_writable(1)), 4, O.CMP_TEXT, rle=-1, batches=3)
    assert exp["rle_used"] and st["adjacent_equal_keys"] > 0.5 * len(words)


def test_text_keys_unique_rle_off_auto():
    rng = random.Random(2)
    words = list({"".join(rng.choice("abcdefgh") for _ in range(rng.randint(0, 24))) for _ in range(30000)})
    rng.shuffle(words)
    exp, st = _check_against_oracle(_text_records(words, lambda w: O.long_writable(len(w))), 5, O.CMP_TEXT, rle=-1)
    assert not exp["rle_used"]


@pytest.mark.parametrize("rle", [0, 1])
def test_prefix_chains_zero_bytes_and_empty_keys(rle):
    """Keys that are prefixes of each other, contain NUL bytes, or are empty: shorter-prefix-first order
    (FastByteComparisons.java:92-116) and the writer's "never RLE an empty key" rule (IFile.java:543)."""
    keys = [b"", b"", b"\x00", b"\x00\x00", b"\x00\x00\x00", b"\x00\x00\x00\x00", b"\x00\x00\x00\x00\x00",
            b"ab", b"ab\x00", b"ab\x00\x00", b"ab\x00\x00\x00\x01", b"ab\x00\x00\x00\x00\x00\x00\x00", b"abc",
            b"\xff", b"\xff\xff\xff\xff\xff", b"\x80", b"\x7f", b"ab", b"ab", b"abc", b"\x00"]
    rng = random.Random(3)
    keys = keys * 3
    rng.shuffle(keys)
    recs = [(k, b"v" + bytes([len(k)]) * (len(k) % 5)) for k in keys]
    for P in (1, 3):
        part = [(len(k) * 7 + (k[0] if k else 0)) % P for k in keys]
        _check_against_oracle(recs, P, O.CMP_BYTES, rle=rle, partition=part)


@pytest.mark.parametrize("rle", [0, 1])
def test_long_shared_prefixes_need_deep_refinement(rle):
    rng = random.Random(4)
    base = bytes(rng.getrandbits(8) for _ in range(40))
    keys = []
    for i in range(3000):
        cut = rng.randint(0, 40)
        keys.append(base[:cut] + bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 3))))
    recs = [(k, zlib.crc32(k).to_bytes(4, "big")) for k in keys]  # value = f(key): ties are byte-identical
    _check_against_oracle(recs, 2, O.CMP_BYTES, rle=rle)


def test_all_keys_identical():
    recs = [(b"samekey!", b"val")] * 5000
    exp, st = _check_against_oracle(recs, 3, O.CMP_BYTES, rle=-1)
    assert exp["rle_used"] and st["adjacent_equal_keys"] == 4999


def test_int_and_long_comparators_signed_order():
    rng = random.Random(6)
    vals = [rng.randint(-2 ** 31, 2 ** 31 - 1) for _ in range(5000)] + [0, -1, 1, -2 ** 31, 2 ** 31 - 1] * 3
    recs = [(O.int_writable(v), O.text(str(v))) for v in vals]
    _check_against_oracle(recs, 4, O.CMP_INT, rle=0)
    lv = [rng.randint(-2 ** 63, 2 ** 63 - 1) for _ in range(3000)] + [0, -1, 1]
    recs = [(O.long_writable(v), b"x") for v in lv]
    _check_against_oracle(recs, 4, O.CMP_LONG, rle=0)


def test_byteswritable_default_serialization():
    rng = random.Random(7)
    recs = []
    for _ in range(4000):
        k = bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 12)))
        recs.append((len(k).to_bytes(4, "big") + k, b"v" * (len(k) % 3)))
    _check_against_oracle(recs, 6, O.CMP_BYTESWRITABLE, rle=0)


def test_many_partitions_mostly_empty_and_empty_partition_details_off():
    rng = random.Random(8)
    recs = [(bytes(rng.getrandbits(8) for _ in range(8)), b"12345678") for _ in range(300)]
    _check_against_oracle(recs, 1000, O.CMP_BYTES, rle=0)
    _check_against_oracle(recs, 1000, O.CMP_BYTES, rle=0, send_empty=False)
    _check_against_oracle([], 5, O.CMP_BYTES, rle=0, send_empty=False)
    _check_against_oracle([], 5, O.CMP_BYTES, rle=0, send_empty=True)


def test_large_values_multi_piece_tiles():
    """4 KB values (BASELINE config 5 record shape): tiles span several shared-memory pieces."""
    rng = np.random.default_rng(9)
    recs = []
    for i in range(600):
        k = int(rng.integers(0, 50)).to_bytes(16, "big")      # Zipf-ish: few distinct keys
        v = (zlib.crc32(k).to_bytes(4, "big") * 1024)[:4096]   # value = f(key)
        recs.append((k, v))
    _check_against_oracle(recs, 4, O.CMP_BYTES, rle=1)
    _check_against_oracle(recs, 4, O.CMP_BYTES, rle=0)


def test_default_sorter_mode_never_rles_single_spill():
    recs = [(O.text("a"), O.text("1")), (O.text("a"), O.text("1"))] * 50
    exp, st = _check_against_oracle(recs, 2, O.CMP_TEXT, rle=-1, legacy=True)
    assert not exp["rle_used"]


def test_illegal_partition_is_an_error():
    recs = [(b"k1", b"v"), (b"k2", b"v")]
    kv, ko, vo, vl = _pack(recs)
    with T.GpuSorter(2, partitioner=T.PART_GIVEN) as s:
        s.collect(kv, ko, vo, vl, np.array([0, 2], np.int32))
        with pytest.raises(IOError, match="Illegal partition"):
            s.flush_to_memory()
    with T.GpuSorter(2, partitioner=T.PART_GIVEN) as s:
        with pytest.raises(IOError):
            s.collect(kv, ko, vo, vl, None)


def test_crc_of_every_segment_matches_zlib():
    n, P = 200000, 16
    kv = O.gen_c2(0, n, seed=3)
    with T.GpuSorter(P, fixed=(16, 64)) as s:
        s.collect_fixed(kv)
        out, _, index, _ = s.flush_to_memory()
    out = bytes(out)
    for p in range(P):
        start, raw, part = (int(x) for x in index[p])
        seg = out[start:start + part]
        assert seg[:4] == b"TIF\x00" and raw == part - 4
        assert int.from_bytes(seg[-4:], "big") == zlib.crc32(seg[4:-4])


@pytest.mark.parametrize("kl,vl", [(8, 8), (16, 16), (4, 12), (16, 32), (24, 72), (16, 112)])
@pytest.mark.parametrize("n", [1, 255, 70001])
def test_fast_emit_other_16_byte_strides(kl, vl, n):
    """The source-oriented emit kernel serves every fixed-width record whose stride is a multiple of 16 bytes
    (1..8 pieces per record, tiles of fewer than 256 records for the wide ones)."""
    rng = np.random.default_rng(kl * 1000 + vl + n)
    P = 5
    kv = rng.integers(0, 256, size=n * (kl + vl), dtype=np.uint8)
    ko = np.arange(n, dtype=np.uint64) * (kl + vl)
    exp = O.pipelined_sort(O.sorter_conf(P), kv, ko, np.full(n, kl, np.uint32), np.full(n, vl, np.uint32))
    with T.GpuSorter(P, fixed=(kl, vl)) as s:
        s.collect_fixed(kv)
        out, index_bytes, _, st = s.flush_to_memory()
    assert bytes(out) == exp["file_out"] and index_bytes == exp["index_out"]


def test_sorter_handle_reset_reuses_allocations():
    with T.GpuSorter(4, fixed=(16, 64)) as s:
        for seed, n in ((1, 5000), (2, 12000), (3, 100)):
            kv = O.gen_c2(0, n, seed=seed)
            s.reset()
            s.collect_fixed(kv)
            out, index_bytes, _, _ = s.flush_to_memory()
            exp = O.pipelined_sort_fixed(O.sorter_conf(4), kv, 16, 64)
            assert bytes(out) == exp["file_out"] and index_bytes == exp["index_out"]


def test_sort_memory_budget_is_enforced_and_spill_recovers():
    """tezgpu_conf.mem_budget_bytes (the memory ExternalSorter was granted): collecting past it is TEZGPU_E_NOMEM, the
    caller spills (flush + reset) and goes on -- the host layer's multi-spill path rests on exactly this."""
    kv = O.gen_c2(0, 3000, seed=9)
    with T.GpuSorter(4, fixed=(16, 64), mem_budget=2000 * 80) as s:
        s.collect_fixed(kv[:1500 * 80])
        with pytest.raises(IOError) as ei:
            s.collect_fixed(kv[1500 * 80:])
        assert ei.value.code == T.E_NOMEM and "budget" in str(ei.value)
        out1, _, _, _ = s.flush_to_memory()
        s.reset()
        s.collect_fixed(kv[1500 * 80:])
        out2, _, _, _ = s.flush_to_memory()
    assert bytes(out1) == O.pipelined_sort_fixed(O.sorter_conf(4), kv[:1500 * 80], 16, 64)["file_out"]
    assert bytes(out2) == O.pipelined_sort_fixed(O.sorter_conf(4), kv[1500 * 80:], 16, 64)["file_out"]
    recs = [(b"k%04d" % i, b"v" * 50) for i in range(100)]
    kvb, ko, vo, vl = _pack(recs)
    with T.GpuSorter(2, mem_budget=3000) as s:
        s.collect(kvb[:int(vo[49] + vl[49])], ko[:50], vo[:50], vl[:50])
        with pytest.raises(IOError, match="budget"):
            s.collect(kvb[int(ko[50]):], ko[50:] - ko[50], vo[50:] - ko[50], vl[50:])
    with T.GpuSorter(2, mem_budget=10) as s:           # a first batch larger than the whole budget is still taken
        s.collect(kvb, ko, vo, vl)
        assert s.flush_to_memory()[3]["output_records"] == 100
