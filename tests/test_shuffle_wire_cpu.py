"""CPU: the ShuffleHandler <-> FetcherOrderedGrouped wire format -- ShuffleHeader framing
(OG/ShuffleHeader.java:76-106) against an independent restatement with the oracle's vint / Text encoders, and the
response splitter on hand-built bodies.  No device needed: the library's host code."""
import random

import pytest

from oracle import tez_oracle as O
from tez_b200 import native
from tez_b200._lib import TezGpuError


def _expected_header(map_id, part_len, raw_len, reduce):
    mid = map_id.encode("utf-8")
    # Text.writeString = vint(byte length) + bytes; then writeVLong x2, writeVInt
    return O.vint(len(mid)) + mid + O.vint(part_len) + O.vint(raw_len) + O.vint(reduce)


@pytest.mark.parametrize("map_id,part,raw,red", [
    ("attempt_1_0001_1_00_000000_0_10003", 10, 6, 0),
    ("attempt_1400000000000_0001_1_00_000123_0_10003_src_17_spill_2", 82000010, 82000006, 63),
    ("m", 127, 128, 1023),
    ("attempt_x", 1 << 31, (1 << 40) + 5, 1 << 20),
    ("", 0, 0, 0),
])
def test_header_bytes_and_round_trip(map_id, part, raw, red):
    h = native.shuffle_header(map_id, part, raw, red)
    assert h == _expected_header(map_id, part, raw, red)
    assert native.read_shuffle_header(h + b"\x99" * 7) == (map_id, part, raw, red, len(h))


def test_header_sizes_follow_hadoop_vint_sizes():
    rng = random.Random(9)
    for _ in range(500):
        part, raw, red = rng.getrandbits(rng.randint(1, 47)), rng.getrandbits(rng.randint(1, 47)), rng.getrandbits(rng.randint(0, 20))
        mid = "attempt_%d" % rng.getrandbits(60)
        assert native.shuffle_header(mid, part, raw, red) == _expected_header(mid, part, raw, red)


def test_malformed_headers_are_rejected():
    good = native.shuffle_header("attempt_1", 100, 96, 3)
    for cut in range(len(good)):
        with pytest.raises(TezGpuError):
            native.read_shuffle_header(good[:cut])
    # map id longer than MAX_ID_LENGTH = 1000 (ShuffleHeader.java:48: readStringSafely)
    with pytest.raises(TezGpuError):
        native.read_shuffle_header(O.vint(1001) + b"a" * 1001 + O.vint(1) + O.vint(1) + O.vint(0))
    # negative lengths
    with pytest.raises(TezGpuError):
        native.read_shuffle_header(O.vint(1) + b"a" + O.vint(-5) + O.vint(1) + O.vint(0))


def test_receive_splits_a_multi_segment_response():
    segs = [O.write_ifile([(O.text("k%d" % i), O.text("v" * i))])[0] for i in range(5)]
    body = b""
    for i, s in enumerate(segs):
        body += native.shuffle_header("attempt_%d" % i, len(s), len(s) - 4, 7) + s
    got = native.shuffle_receive(body)
    assert [(m, r, raw, seg) for m, r, raw, seg in got] == [("attempt_%d" % i, 7, len(s) - 4, s) for i, s in enumerate(segs)]
    with pytest.raises(TezGpuError):
        native.shuffle_receive(body[:-3])     # the response ends inside the last segment
    assert native.shuffle_receive(b"") == []
    many = b"".join(native.shuffle_header("a%d" % i, 0, 0, i) for i in range(200))   # more segments than the first table holds
    assert [r for _, r, _, _ in native.shuffle_receive(many)] == list(range(200))
