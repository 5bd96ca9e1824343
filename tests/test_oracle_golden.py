"""Pins the CPU oracle (oracle/) against every golden vector / known-answer test the reference holds for the
hot path.  CPU only.

RLT/ = /root/reference/tez-runtime-library/src/test/java/org/apache/tez/runtime/library/
"""
import os
import random
import zlib

import numpy as np
import pytest

from oracle import tez_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# RLT/common/sort/impl/TestIFile.java:395-397
FIXTURE_RAWS = [2392, 102314, 42576, 31432, 25090]
FIXTURE_COMPRESSED = [723, 25396, 10926, 8203, 6665]


def _fixture_segments():
    data = open(os.path.join(GOLDEN, "TestIFile_concatenated_compressed.bin"), "rb").read()
    assert len(data) == sum(FIXTURE_COMPRESSED) == 51913
    segs, pos = [], 0
    for c in FIXTURE_COMPRESSED:
        segs.append(data[pos:pos + c])
        pos += c
    return segs


def test_fixture_header_crc_and_lengths():
    """Header, CRC scope (body only, big-endian trailer) and rawLength = 4 + body, as written by the real IFile.Writer."""
    for seg, raw in zip(_fixture_segments(), FIXTURE_RAWS):
        assert seg[:4] == b"TIF\x01"
        assert int.from_bytes(seg[-4:], "big") == zlib.crc32(seg[4:-4]) == O.crc32(seg[4:-4])
        body = zlib.decompress(seg[4:-4])
        assert len(body) == raw - 4
        assert body.endswith(b"\xff\xff")


def test_fixture_records_parse_and_reencode_bit_exact():
    """The oracle reader walks every record of the reference-written bodies, and the oracle writer reproduces the
    exact body bytes from the parsed records (vint framing, record layout, EOF markers)."""
    first = True
    total = 0
    for seg in _fixture_segments():
        body = zlib.decompress(seg[4:-4])
        inmem = body + zlib.crc32(body).to_bytes(4, "big")  # in-memory segment = body + crc (OG/InMemoryWriter.java:55-77)
        recs = O.read_ifile(inmem, has_header=False)
        assert recs, "fixture segment has records"
        for ks, k, v in recs:
            assert ks == O.NEW_KEY
            assert len(v) == 4                      # IntWritable
            assert k[0] == len(k) - 1               # Text: vint(len) + utf8 (short keys)
        if first:
            assert recs[0][1] == b"\x04key0" and recs[0][2] == b"\x00\x00\x00\x00"
            assert body[:11] == bytes.fromhex("0504046b65793000000000")
            first = False
        total += len(recs)
        out, raw_len, part_len = O.write_ifile([(k, v) for _, k, v in recs], rle=False)
        assert out[:4] == b"TIF\x00"
        assert out[4:-4] == body
        assert raw_len == len(body) + 4
        assert part_len == len(out) == raw_len + 4
        assert int.from_bytes(out[-4:], "big") == zlib.crc32(body)
    assert total > 1000


def test_empty_segment_bytes():
    """RLT/common/sort/impl/TestPipelinedSorter.java:239-243, dflt/TestDefaultSorter.java:343-352: 6 raw / 10 file bytes."""
    out, raw_len, part_len = O.write_ifile([])
    assert out == bytes.fromhex("54494600ffffffff0000")
    assert (raw_len, part_len) == (6, 10)


@pytest.mark.parametrize("v,enc", [
    (0, "00"), (127, "7f"), (-112, "90"), (128, "8f80"), (255, "8fff"), (256, "8e0100"), (4096, "8e1000"),
    (-1, "ff"), (-2, "fe"), (-3, "fd"), (-113, "8770"), (65536, "8d010000"), (2 ** 31 - 1, "8c7fffffff"),
])
def test_vint_known_answers(v, enc):
    b = O.vint(v)
    assert b.hex() == enc
    assert O.lib().tzo_vint_size(v) == len(b)
    assert O.read_vint(b) == (v, len(b))
    assert O.lib().tzo_decode_vint_size(b[0]) == len(b)


def test_vint_roundtrip_random():
    rng = random.Random(7)
    for _ in range(2000):
        v = rng.choice([rng.randint(-200, 200), rng.randint(-2 ** 31, 2 ** 31 - 1), rng.randint(-2 ** 62, 2 ** 62)])
        b = O.vint(v)
        assert O.read_vint(b) == (v, len(b))


def test_crc32_matches_zlib():
    rng = random.Random(3)
    for n in [0, 1, 7, 8, 9, 63, 64, 1000, 65537]:
        d = bytes(rng.getrandbits(8) for _ in range(n))
        assert O.crc32(d) == zlib.crc32(d)
    a, b = b"hello ", b"world"
    assert O.crc32(b, O.crc32(a)) == zlib.crc32(a + b)


# ------------------------------------------------------------------ TestDefaultSorter.testEmptyCaseFileLengths
def _text_records(keys, values):
    kv = bytearray()
    ko, kl, vl = [], [], []
    for k, v in zip(keys, values):
        kb, vb = O.text(k), O.text(v)
        ko.append(len(kv))
        kl.append(len(kb))
        vl.append(len(vb))
        kv += kb + vb
    return bytes(kv), ko, kl, vl


@pytest.mark.parametrize("P,keys,values", [
    (50, ["a", "b"], ["1", "2"]),
    (50, ["a", "a"], ["1", "2"]),
    (50, ["aaa", "bbb", "aaa"], ["1", "2", "3"]),
    (1, ["abcdefghij"], ["1234567890"]),
])
def test_default_sorter_empty_case_file_lengths(P, keys, values):
    """RLT/common/sort/impl/dflt/TestDefaultSorter.java:295-356: file.out length =
    10 * nonEmptyPartitions + sum(len(k)+2 + len(v)+2); numSpills == 1."""
    kv, ko, kl, vl = _text_records(keys, values)
    conf = O.sorter_conf(P, cmp_kind=O.CMP_TEXT, partitioner=O.PART_HASH, legacy=True)
    r = O.pipelined_sort(conf, kv, ko, kl, vl)
    non_empty = len({O.partition_of(O.CMP_TEXT, O.text(k), P) for k in keys})
    expected = non_empty * 10 + sum(len(k) + 2 + len(v) + 2 for k, v in zip(keys, values))
    assert len(r["file_out"]) == expected
    idx = r["index"]
    assert int((idx[:, 1] > 6).sum()) == non_empty          # TezIndexRecord.hasData: rawLength > 6
    assert int((idx[:, 1] == 0).sum()) == P - non_empty      # (pos,0,0) entries for empty partitions
    assert len(r["index_out"]) == P * 24 + 8
    # counters identities of RLT/.../TestPipelinedSorter.java:593-626
    c = r["counters"]
    assert c["OUTPUT_RECORDS"] == len(keys) == c["SPILLED_RECORDS"]
    assert c["OUTPUT_BYTES_PHYSICAL"] == expected
    assert c["OUTPUT_BYTES_WITH_OVERHEAD"] == int(idx[:, 1].sum())


def test_pipelined_sorter_rle_when_all_keys_equal():
    """PipelinedSorter with {a,a}: one compare returning 0 => eq=1 > 0.1*2 => RLE (SORT/PipelinedSorter.java:1436-1438).
    Bytes follow SORT/IFile.java:559-614: full record, then FE vint(vlen) val, FD before EOF."""
    kv, ko, kl, vl = _text_records(["a", "a"], ["1", "2"])
    r = O.pipelined_sort(O.sorter_conf(1, cmp_kind=O.CMP_TEXT), kv, ko, kl, vl)
    assert r["rle_used"] and r["eq"] == 1
    assert r["file_out"][4:-4].hex() == "0202" + "0161" + "0131" + "fe" + "02" + "0132" + "fd" + "ffff"


def test_send_empty_partition_details_off_writes_empty_segments():
    kv, ko, kl, vl = _text_records(["a"], ["1"])
    r = O.pipelined_sort(O.sorter_conf(4, cmp_kind=O.CMP_TEXT, send_empty=False), kv, ko, kl, vl)
    assert len(r["file_out"]) == 3 * 10 + (10 + 6)
    assert (r["index"][:, 1] >= 6).all()


# ------------------------------------------------------------------ TestTezMerger literal tables
D, S = False, True
MERGER_TABLES = {
    # name: (files, expected [(key, same)]) -- RLT/common/sort/impl/TestTezMerger.java
    "WithEmptyStrings:185-233": ([["0", "0"], [""], ["0", "0"], ["1", "2"]],
                                 [("", D), ("0", D), ("0", S), ("0", S), ("0", S), ("1", D), ("2", D)]),
    "No_RLE:235-283": ([["1", "4", "5"], ["2", "6", "7"], ["3", "8", "9"]],
                       [(str(i), D) for i in range(1, 10)]),
    "RLE_acrossFiles:285-322": ([["0", "0"], ["0", "1"]], [("0", D), ("0", S), ("0", S), ("1", D)]),
    "mixedFiles:324-369": ([["0"], [], [""], ["0", "0", "0"]], [("", D), ("0", D), ("0", S), ("0", S), ("0", S)]),
    "RLE:371-406": ([["1", "2", "2"], ["2", "3"]], [("1", D), ("2", D), ("2", S), ("2", S), ("3", D)]),
    "RLE2:408-454": ([["0", "1", "1"], ["0", "1"], ["0", "1", "1"]],
                     [("0", D), ("0", S), ("0", S), ("1", D), ("1", S), ("1", S), ("1", S), ("1", S)]),
    "plain:456-493": ([["0"], ["0"], ["1"]], [("0", D), ("0", S), ("1", D)]),
    "RLE3:495-526": ([["0"], ["0", "1", "1"]], [("0", D), ("0", S), ("1", D), ("1", S)]),
    "allEmptyFiles:528-552": ([[], [], [], []], []),
}


def _ifile_with_text_data(keys, salt):
    # createIFileWithTextData :140-151 -- Writer(rle=true), value = key + "_" + nanoTime
    recs = [(O.text(k), O.text("%s_%d" % (k, salt * 1000 + i))) for i, k in enumerate(keys)]
    return O.write_ifile(recs, rle=True)[0]


@pytest.mark.parametrize("name", sorted(MERGER_TABLES))
def test_tez_merger_known_answer_tables(name):
    files, expected = MERGER_TABLES[name]
    segs = [_ifile_with_text_data(keys, i) for i, keys in enumerate(files)]
    # merge(pathList, rc) :557-566 -- factor 4, Path[] ctor sorts segments by length and sets checkForSameKeys
    r = O.merge(segs, O.CMP_SIGNED_BYTES, factor=4, sort_segments=True, check_for_same_keys=True)
    got = [(k[1:].decode(), same) for k, _, same in r["records"]]
    assert got == expected
    # every input record is emitted exactly once
    assert sorted(v for _, v, _ in r["records"]) == sorted(v for s in segs for _, _, v in O.read_ifile(s))


def test_merger_writefile_emits_rle_markers_for_same_keys():
    """TezMerger.writeFile (SORT/TezMerger.java:215-245) passes REPEAT_KEY for isSameKey() records even when the
    writer itself was created without rle."""
    segs = [_ifile_with_text_data(["0", "1"], 1), _ifile_with_text_data(["0", "1"], 2)]
    r = O.merge(segs, O.CMP_TEXT, factor=10, writer_rle=False)
    states = [ks for ks, _, _ in O.read_ifile(r["ifile"])]
    assert states == [O.NEW_KEY, O.SAME_KEY, O.NEW_KEY, O.SAME_KEY]
    assert r["comp_len"] == len(r["ifile"]) == r["raw_len"] + 4


@pytest.mark.parametrize("nfiles,keys_per_file,factor", [(1, 10, 2), (5, 100, 3), (20, 50, 4), (57, 11, 5), (3, 0, 2)])
def test_merger_multi_pass_matches_sorted(nfiles, keys_per_file, factor):
    """testMerge grid (RLT/.../TestTezMerger.java:110-154): IntWritable keys, LongWritable values, multi-pass by factor."""
    rng = random.Random(nfiles * 1000 + keys_per_file)
    segs, everything = [], []
    for f in range(nfiles):
        keys = sorted(rng.randint(-1000, 1000) for _ in range(keys_per_file))
        recs = [(O.int_writable(k), O.long_writable(rng.getrandbits(40))) for k in keys]
        everything += recs
        segs.append(O.write_ifile(recs, rle=(f % 2 == 0))[0])
    r = O.merge(segs, O.CMP_INT, factor=factor, sort_segments=True)
    keys = [int.from_bytes(k, "big", signed=True) for k, _, _ in r["records"]]
    assert keys == sorted(keys)
    assert sorted((k, v) for k, v, _ in r["records"]) == sorted(everything)
    if nfiles > factor and keys_per_file > 0:
        assert r["passes"] >= 1
    for i, (k, _, same) in enumerate(r["records"]):
        if same:
            assert i > 0 and r["records"][i - 1][0] == k
    back = O.read_ifile(r["ifile"])
    assert [(k, v) for _, k, v in back] == [(k, v) for k, v, _ in r["records"]]


# ------------------------------------------------------------------ TestProxyComparator
PROXY_KEYS = ["", "A", "B", "AA", "BB", "BA", "CB", "AAA", "BBBB", "CCCCC",
              "æAAAA", "÷", "A÷", "÷AAAAAAAAA", "÷÷",
              "÷÷ææA", "÷÷ææA"]


def test_proxy_comparator_table():
    """RLT/common/comparator/TestProxyComparator.java:36-82: proxy order implies comparator order."""
    for l in PROXY_KEYS:
        for r in PROXY_KEYS:
            lb, rb = l.encode("utf-8"), r.encode("utf-8")
            lp, rp = O.proxy(lb), O.proxy(rb)
            c = O.compare(O.CMP_BYTES, lb, rb)
            if lp < rp:
                assert c < 0
            if lp > rp:
                assert c > 0
            assert (c == 0) == (lb == rb)
            assert (c < 0) == (lb < rb)  # python bytes order == unsigned lexicographic, shorter prefix first


def test_pipelined_prefix_word():
    """P=64 => partitionBits=8 => prefix = p<<24 | k0<<8 | k1 ; sign bit always clear."""
    key = bytes([0xAB, 0xCD, 0xEF, 0x01])
    assert O.lib().tzo_pipelined_prefix(63, O.proxy(key), 64) == (63 << 24) | (0xAB << 8) | 0xCD
    assert O.lib().tzo_pipelined_prefix(1023, O.proxy(key), 1024) == (1023 << 20) | (0xABCDEF >> 12)


# ------------------------------------------------------------------ misc restatement sanity
def test_hash_bytes_definition():
    """WritableComparator.hashBytes (parity UNPINNED by reference tests): h=1; h=31*h+(signed)b."""
    assert O.hash_bytes(b"") == 1
    assert O.hash_bytes(b"a") == 31 + 97
    assert O.hash_bytes(bytes([0xFF])) == 31 - 1
    h = 1
    for b in b"hello world":
        h = (31 * h + b) & 0xFFFFFFFF
    assert O.hash_bytes(b"hello world") & 0xFFFFFFFF == h
    assert O.key_hash(O.CMP_TEXT, O.text("hello world")) == O.hash_bytes(b"hello world")
    assert O.key_hash(O.CMP_INT, O.int_writable(-5)) == -5


def test_spill_record_layout():
    idx = [0, 100, 104, 104, 0, 0, 104, 50, 54]
    b = O.spill_record_bytes(idx)
    assert len(b) == 3 * 24 + 8
    assert [int.from_bytes(b[i * 8:i * 8 + 8], "big") for i in range(9)] == idx
    assert int.from_bytes(b[-8:], "big") == zlib.crc32(b[:-8])


def test_rle_marker_roundtrip():
    """RLT/common/sort/impl/TestIFile.java:299-374 (testWithRLEMarker): repeated keys come back as SAME_KEY."""
    recs = [(O.text("k1"), O.int_writable(1)), (O.text("k1"), O.int_writable(2)), (O.text("k1"), O.int_writable(3)),
            (O.text("k2"), O.int_writable(4)), (O.text("k3"), O.int_writable(5)), (O.text("k3"), O.int_writable(6))]
    seg, raw, part = O.write_ifile(recs, rle=True)
    back = O.read_ifile(seg)
    assert [ks for ks, _, _ in back] == [O.NEW_KEY, O.SAME_KEY, O.SAME_KEY, O.NEW_KEY, O.NEW_KEY, O.SAME_KEY]
    assert [(k, v) for _, k, v in back] == recs
    plain = O.write_ifile(recs, rle=False)[0]
    assert len(seg) < len(plain)
    # explicit REPEAT_KEY (writer without rle) produces identical bytes
    rep = [(recs[0][0], recs[0][1]), (None, recs[1][1]), (None, recs[2][1]), recs[3], recs[4], (None, recs[5][1])]
    assert O.write_ifile(rep, rle=False)[0] == seg


def test_pipelined_sort_random_vs_python_sorted():
    rng = np.random.default_rng(11)
    n, P = 5000, 7
    kv = O.gen_c2(0, n, seed=2)
    r = O.pipelined_sort_fixed(O.sorter_conf(P, span_records=512), kv, 16, 64)
    rec = kv.reshape(n, 80)
    parts = np.array([O.partition_of(O.CMP_BYTES, rec[i, :16].tobytes(), P) for i in range(n)])
    out = bytearray()
    for p in range(P):
        rows = sorted((rec[i].tobytes() for i in np.nonzero(parts == p)[0]), key=lambda b: b[:16])
        seg = O.write_ifile([(b[:16], b[16:]) for b in rows])[0]
        assert int(r["index"][p, 0]) == len(out)
        out += seg
    assert bytes(out) == r["file_out"]
    assert not r["rle_used"] and r["eq"] == 0
    del rng


def test_ordered_word_count_known_answer_through_the_oracle():
    """TestTezJobs.generateOrderedWordCountInput / verifyOutput (tez-tests/.../TestTezJobs.java:748-805):
    words a_1..a_10 with counts 20,18,..,2.  Three tokenizer tasks sort (Text word, IntWritable 1) into 4 partitions,
    each summation task merges its partition of every producer and groups equal keys, a second ordered edge sorts
    (IntWritable count, Text word) -- the whole map-sort / merge / group path on the CPU restatement."""
    import random
    words = []
    for i in range(1, 11):
        words += ["a_%d" % i] * (22 - 2 * i)
    random.Random(3).shuffle(words)
    P = 4

    def sort_records(records, cmp_kind, nparts):
        kv = bytearray()
        ko, kl, vl = [], [], []
        for k, v in records:
            ko.append(len(kv)); kl.append(len(k)); vl.append(len(v))
            kv += k + v
        return O.pipelined_sort(O.sorter_conf(nparts, cmp_kind=cmp_kind), np.frombuffer(bytes(kv), np.uint8),
                                np.array(ko, np.uint64), np.array(kl, np.uint32), np.array(vl, np.uint32))

    producers = [sort_records([(O.text(w), O.int_writable(1)) for w in words[t::3]], O.CMP_TEXT, P) for t in range(3)]
    counts, seen = {}, 0
    for p in range(P):
        runs = []
        for r in producers:
            start, _, part = (int(x) for x in r["index"][p])
            if part:
                runs.append(r["file_out"][start:start + part])
        if not runs:
            continue
        merged = O.merge(runs, O.CMP_TEXT, factor=100)
        prev = None
        for key, val, same in merged["records"]:
            word = key[1:].decode()
            if prev is None or key != prev:
                assert word not in counts          # a key's records are contiguous: ValuesIterator grouping
                counts[word] = 0
            assert same == (prev is not None and key == prev)
            counts[word] += int.from_bytes(val, "big")
            prev = key
            seen += 1
    assert seen == len(words)
    assert counts == {"a_%d" % i: 22 - 2 * i for i in range(1, 11)}
    # every word went to the partition HashPartitioner assigns to its Text key
    for r in producers:
        for p in range(P):
            start, _, part = (int(x) for x in r["index"][p])
            for _, key, _ in O.read_ifile(r["file_out"][start:start + part]) if part else []:
                assert O.partition_of(O.CMP_TEXT, key, P) == p
    second = sort_records([(O.int_writable(c), O.text(w)) for w, c in counts.items()], O.CMP_INT, 1)
    final = [(int.from_bytes(k, "big", signed=True), v[1:].decode()) for _, k, v in O.read_ifile(second["file_out"])]
    assert final == [(22 - 2 * i, "a_%d" % i) for i in range(10, 0, -1)]
