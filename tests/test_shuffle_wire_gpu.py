"""GPU: SURVEY 8 f-2 -- map outputs that live in HBM served through the ShuffleHandler wire format and consumed like
FetcherOrderedGrouped.copyMapOutput does (header, reserve, read partLength bytes), then merged on the device: the result
must equal TezMerger over the producers' segments, and every served segment must be the producer's file.out bytes."""
import numpy as np
import pytest
import torch

from oracle import tez_oracle as O
import tez_b200 as T
from tez_b200 import native

pytestmark = pytest.mark.gpu


def test_serve_receive_merge_matches_oracle():
    P, G, n = 8, 3, 20000
    bodies, files = [], []
    for g in range(G):
        kv = O.gen_c2(g * n, n, seed=31)
        exp = O.pipelined_sort_fixed(O.sorter_conf(P), kv, 16, 64)
        with T.GpuSorter(P, fixed=(16, 64)) as s:
            d_kv = torch.from_numpy(kv).cuda()
            cap = n * 82 + 16 * P + 4096
            d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
            out_len, index, st = s.sort_device_fixed(d_kv.data_ptr(), n, d_out.data_ptr(), cap)
            assert out_len == len(exp["file_out"]) and np.array_equal(index, exp["index"])
            # reducers 2..5 of this map output, as one response body
            bodies.append(native.shuffle_serve(d_out.data_ptr(), index, "attempt_1_0001_1_00_%06d_0" % g, 2, 4))
            files.append((exp["file_out"], exp["index"]))
    for p in range(2, 6):
        segs = []
        for g in range(G):
            got = native.shuffle_receive(bodies[g])
            assert [r for _, r, _, _ in got] == [2, 3, 4, 5]
            mid, red, raw, seg = got[p - 2]
            start, raw_len, part_len = (int(x) for x in files[g][1][p])
            assert mid == "attempt_1_0001_1_00_%06d_0" % g and red == p and raw == raw_len
            assert seg == files[g][0][start:start + part_len]
            segs.append(seg)
        exp = O.merge(segs, O.CMP_BYTES, factor=100)
        with T.GpuMerger(segs, comparator=T.CMP_BYTES) as m:
            merged, raw, part, _ = m.write_ifile()
        assert merged == exp["ifile"]


def test_serve_rejects_a_short_buffer_and_empty_partitions_travel_as_ten_bytes():
    P, n = 64, 50          # most partitions empty
    kv = O.gen_c2(0, n, seed=33)
    with T.GpuSorter(P, fixed=(16, 64)) as s:
        d_kv = torch.from_numpy(kv).cuda()
        d_out = torch.empty(n * 82 + 16 * P + 4096, dtype=torch.uint8, device="cuda")
        out_len, index, st = s.sort_device_fixed(d_kv.data_ptr(), n, d_out.data_ptr(), d_out.numel())
        got = native.shuffle_receive(native.shuffle_serve(d_out.data_ptr(), index, "attempt_9", 0, P))
        host = d_out[:out_len].cpu().numpy().tobytes()
    assert len(got) == P
    for p, (mid, red, raw, seg) in enumerate(got):
        start, raw_len, part_len = (int(x) for x in index[p])
        assert red == p and raw == raw_len and seg == host[start:start + part_len]
        if part_len == 10:
            assert raw_len == 6 and O.read_ifile(seg) == []
    import ctypes as C
    from tez_b200 import _lib
    L = _lib.load()
    idx = np.ascontiguousarray(index, dtype=np.int64)
    small = np.empty(16, dtype=np.uint8)
    nn = C.c_uint64()
    rc = L.tezgpu_shuffle_serve(0, d_out.data_ptr(), idx.ctypes.data, b"attempt_9", 0, P, small.ctypes.data, small.size, C.byref(nn), None)
    from tez_b200 import constants
    assert rc == constants.E_NOMEM
