"""Property tests of the CPU oracle (the checker itself): size-independent invariants of the IFile writer/reader, the
PipelinedSorter restatement and the TezMerger restatement on randomly drawn inputs (hypothesis, derandomised)."""
import zlib

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import tez_oracle as O

KEY = st.binary(min_size=0, max_size=12)
VAL = st.binary(min_size=0, max_size=20)
SET = dict(max_examples=60, deadline=None, derandomize=True)


@settings(**SET)
@given(st.lists(st.tuples(KEY, VAL), max_size=40), st.booleans())
def test_ifile_writer_reader_round_trip(records, rle):
    """IFile.Writer -> IFile.Reader (SORT/IFile.java:262-634, 877-1000): records come back in order whatever the RLE
    setting; header, EOF markers and the big-endian CRC32 trailer of the body are in place."""
    seg, raw, part = O.write_ifile(records, rle=rle)
    assert seg[:4] == b"TIF\x00" and part == len(seg) == raw + 4
    assert int.from_bytes(seg[-4:], "big") == zlib.crc32(seg[4:-4])
    got = [(k, v) for _, k, v in O.read_ifile(seg)]
    assert got == records


@settings(**SET)
@given(st.lists(st.tuples(st.binary(min_size=1, max_size=10), VAL), min_size=0, max_size=80), st.integers(1, 9))
def test_pipelined_sort_partitions_are_sorted_and_complete(records, P):
    """Every record lands in the partition HashPartitioner gives its key (RL/partitioner/HashPartitioner.java:33-35),
    each partition segment is ordered by the raw comparator, nothing is lost or invented, and the index triples
    (start, rawLength, partLength) tile file.out exactly."""
    kv = bytearray()
    ko, kl, vl = [], [], []
    for k, v in records:
        ko.append(len(kv)); kl.append(len(k)); vl.append(len(v))
        kv += k + v
    r = O.pipelined_sort(O.sorter_conf(P, rle_policy=0), np.frombuffer(bytes(kv), np.uint8), np.array(ko, np.uint64),
                         np.array(kl, np.uint32), np.array(vl, np.uint32))
    pos, seen = 0, []
    for p in range(P):
        start, raw, part = (int(x) for x in r["index"][p])
        assert start == pos
        pos += part
        if not part:
            assert raw == 0
            continue
        assert part == raw + 4
        recs = [(k, v) for _, k, v in O.read_ifile(r["file_out"][start:start + part])]
        keys = [k for k, _ in recs]
        assert keys == sorted(keys)                      # unsigned lexicographic == TezBytesComparator
        assert all(O.partition_of(O.CMP_BYTES, k, P) == p for k in keys)
        seen += recs
    assert pos == len(r["file_out"])
    assert sorted(seen) == sorted(records)


@settings(**SET)
@given(st.lists(st.lists(st.tuples(st.binary(min_size=1, max_size=6), st.binary(min_size=0, max_size=4)), max_size=25),
                min_size=1, max_size=6), st.integers(2, 5))
def test_merge_is_the_stable_sort_of_the_runs(runs, factor):
    """TezMerger (SORT/TezMerger.java:465-1065): merging sorted runs yields all records in comparator order, whatever
    the merge factor (number of passes); distinct keys keep their values, and isSameKey is true exactly when a key
    repeats the previous one."""
    # keys are made unique inside a run: within one segment the merger learns about a repeated key only from the
    # producer's RLE marker (SORT/TezMerger.java:641-652), so plain equal neighbours of one run are reported DIFF
    runs = [sorted(dict(run).items()) for run in runs]
    segs = [O.write_ifile(run)[0] for run in runs]
    out = O.merge(segs, O.CMP_BYTES, factor=factor)
    keys = [k for k, _, _ in out["records"]]
    assert keys == sorted(keys)
    assert sorted((k, v) for k, v, _ in out["records"]) == sorted(kv for run in runs for kv in run)
    for i, (k, _, same) in enumerate(out["records"]):
        assert same == (i > 0 and keys[i - 1] == k)
    assert [(k, v) for _, k, v in O.read_ifile(out["ifile"])] == [(k, v) for k, v, _ in out["records"]]


def test_repeated_key_inside_one_run_is_not_same_key_without_rle():
    """The observation behind the restriction above, pinned: two equal keys written WITHOUT run-length encoding in a
    single segment come out of the merge flagged DIFF (the writer of the merged file may still RLE them itself)."""
    seg = O.write_ifile([(b"k", b"1"), (b"k", b"2")])[0]
    assert [same for _, _, same in O.merge([seg], O.CMP_BYTES)["records"]] == [False, False]
    seg_rle = O.write_ifile([(b"k", b"1"), (b"k", b"2")], rle=True)[0]
    assert [same for _, _, same in O.merge([seg_rle], O.CMP_BYTES)["records"]] == [False, True]
