"""ctypes binding of the CPU oracle (oracle/libtezoracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg.  The product package
(tez_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtezoracle.so")

CMP_BYTES, CMP_TEXT, CMP_BYTESWRITABLE, CMP_INT, CMP_LONG, CMP_SIGNED_BYTES = range(6)
PART_GIVEN, PART_HASH = 0, 1
NO_KEY, NEW_KEY, SAME_KEY = 0, 1, 2


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("tez_oracle.c", "tez_oracle.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src if os.path.exists(s))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


class Buf(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_size_t), ("cap", C.c_size_t)]

    def bytes(self):
        return C.string_at(self.data, self.len) if self.len else b""


class Writer(C.Structure):
    _fields_ = [("out", C.POINTER(Buf)), ("start", C.c_size_t), ("rle", C.c_int), ("prev_is_repeat", C.c_int),
                ("previous", Buf), ("crc", C.c_uint32), ("raw_len", C.c_int64), ("comp_len", C.c_int64),
                ("records", C.c_int64), ("rle_written", C.c_int64), ("closed", C.c_int)]


class Reader(C.Structure):
    _fields_ = [("data", C.c_void_p), ("pos", C.c_size_t), ("end", C.c_size_t),
                ("cur_klen", C.c_int64), ("cur_vlen", C.c_int64), ("prev_klen", C.c_int64), ("orig_klen", C.c_int64),
                ("key", C.c_void_p), ("klen", C.c_int), ("eof", C.c_int), ("records", C.c_int64), ("length", C.c_size_t)]


class SorterConf(C.Structure):
    _fields_ = [("num_partitions", C.c_int), ("cmp_kind", C.c_int), ("partitioner", C.c_int),
                ("send_empty_partition_details", C.c_int), ("rle_policy", C.c_int), ("span_records", C.c_int),
                ("sort_threads", C.c_int), ("legacy", C.c_int)]


class SorterResult(C.Structure):
    _fields_ = [("file_out", Buf), ("index_out", Buf), ("index", C.POINTER(C.c_int64)), ("rle_used", C.c_int),
                ("eq", C.c_int64), ("total", C.c_int64),
                ("output_records", C.c_int64), ("output_bytes", C.c_int64), ("output_bytes_with_overhead", C.c_int64),
                ("output_bytes_physical", C.c_int64), ("spilled_records", C.c_int64)]


class Segment(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t), ("has_header", C.c_int)]


class MergeResult(C.Structure):
    _fields_ = [("n", C.c_uint64), ("keys", Buf), ("vals", Buf), ("key_len", C.POINTER(C.c_uint32)),
                ("val_len", C.POINTER(C.c_uint32)), ("same_key", C.POINTER(C.c_uint8)), ("ifile", Buf),
                ("raw_len", C.c_int64), ("comp_len", C.c_int64), ("passes", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.tzo_vint_size.argtypes = [C.c_int64]
        L.tzo_write_vlong.argtypes = [C.c_char_p, C.c_int64]
        L.tzo_read_vlong.argtypes = [C.c_char_p, C.POINTER(C.c_int64)]
        L.tzo_crc32.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        L.tzo_crc32.restype = C.c_uint32
        L.tzo_hash_bytes.argtypes = [C.c_char_p, C.c_size_t]
        L.tzo_hash_bytes.restype = C.c_int32
        L.tzo_key_hash.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
        L.tzo_key_hash.restype = C.c_int32
        L.tzo_hash_partition.argtypes = [C.c_int32, C.c_int32]
        L.tzo_hash_partition.restype = C.c_int32
        L.tzo_compare.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.tzo_bytes_proxy.argtypes = [C.c_char_p, C.c_int]
        L.tzo_bytes_proxy.restype = C.c_int32
        L.tzo_pipelined_prefix.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.tzo_pipelined_prefix.restype = C.c_int32
        L.tzo_buf_init.argtypes = [C.POINTER(Buf)]
        L.tzo_buf_free.argtypes = [C.POINTER(Buf)]
        L.tzo_writer_open.argtypes = [C.POINTER(Writer), C.POINTER(Buf), C.c_int]
        L.tzo_writer_append.argtypes = [C.POINTER(Writer), C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.tzo_writer_close.argtypes = [C.POINTER(Writer)]
        L.tzo_reader_open.argtypes = [C.POINTER(Reader), C.c_char_p, C.c_size_t, C.c_int, C.c_int]
        L.tzo_reader_read_raw_key.argtypes = [C.POINTER(Reader)]
        L.tzo_reader_next_raw_value.argtypes = [C.POINTER(Reader), C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        L.tzo_spill_record_bytes.argtypes = [C.POINTER(C.c_int64), C.c_int, C.POINTER(Buf)]
        L.tzo_pipelined_sort.argtypes = [C.POINTER(SorterConf), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_uint64, C.POINTER(SorterResult)]
        L.tzo_pipelined_sort_fixed.argtypes = [C.POINTER(SorterConf), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64,
                                               C.POINTER(SorterResult)]
        L.tzo_sorter_result_free.argtypes = [C.POINTER(SorterResult)]
        L.tzo_merge.argtypes = [C.POINTER(Segment), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(MergeResult)]
        L.tzo_merge_result_free.argtypes = [C.POINTER(MergeResult)]
        L.tzo_splitmix64.argtypes = [C.c_uint64]
        L.tzo_splitmix64.restype = C.c_uint64
        L.tzo_gen_c2.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
        L.tzo_bench_pipelined_fixed.argtypes = [C.POINTER(SorterConf), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64,
                                                C.c_int, C.POINTER(C.c_uint64)]
        L.tzo_bench_pipelined_fixed.restype = C.c_double
        _lib = L
    return _lib


# ---------------------------------------------------------------- small helpers
def vint(v):
    b = C.create_string_buffer(10)
    n = lib().tzo_write_vlong(b, v)
    return b.raw[:n]


def read_vint(data, pos=0):
    out = C.c_int64()
    n = lib().tzo_read_vlong(data[pos:pos + 10].ljust(10, b"\0"), C.byref(out))
    return out.value, n


def crc32(data, crc=0):
    return lib().tzo_crc32(crc, data, len(data))


def hash_bytes(data):
    return lib().tzo_hash_bytes(data, len(data))


def key_hash(kind, key):
    return lib().tzo_key_hash(kind, key, len(key))


def partition_of(kind, key, num_partitions):
    return lib().tzo_hash_partition(key_hash(kind, key), num_partitions)


def compare(kind, a, b):
    return lib().tzo_compare(kind, a, len(a), b, len(b))


def proxy(content):
    return lib().tzo_bytes_proxy(content, len(content))


def text(s):
    """hadoop Text serialisation: vint(byteLen) + utf-8."""
    b = s.encode("utf-8") if isinstance(s, str) else bytes(s)
    return vint(len(b)) + b


def int_writable(v):
    return int(v & 0xFFFFFFFF).to_bytes(4, "big")


def long_writable(v):
    return int(v & 0xFFFFFFFFFFFFFFFF).to_bytes(8, "big")


def write_ifile(records, rle=False):
    """records: iterable of (key_bytes_or_None_for_REPEAT_KEY, value_bytes). Returns (segment bytes, rawLen, partLen)."""
    L = lib()
    out = Buf()
    L.tzo_buf_init(C.byref(out))
    w = Writer()
    L.tzo_writer_open(C.byref(w), C.byref(out), 1 if rle else 0)
    for k, v in records:
        if k is None:
            L.tzo_writer_append(C.byref(w), None, 0, v, len(v))
        else:
            L.tzo_writer_append(C.byref(w), k if len(k) else b"\0", len(k), v, len(v))
    L.tzo_writer_close(C.byref(w))
    data = out.bytes()
    L.tzo_buf_free(C.byref(out))
    return data, w.raw_len, w.comp_len


def read_ifile(seg, has_header=True, verify_crc=True):
    """Returns list of (key_state, key, value) following IFile.Reader semantics."""
    L = lib()
    r = Reader()
    rc = L.tzo_reader_open(C.byref(r), seg, len(seg), 1 if has_header else 0, 1 if verify_crc else 0)
    if rc != 0:
        raise IOError("tzo_reader_open rc=%d" % rc)
    base = r.data
    out = []
    while True:
        ks = L.tzo_reader_read_raw_key(C.byref(r))
        if ks < 0:
            raise IOError("malformed IFile segment")
        if ks == NO_KEY:
            break
        koff = r.key - base
        key = seg[koff:koff + r.klen]
        vp = C.c_void_p()
        vl = C.c_int()
        L.tzo_reader_next_raw_value(C.byref(r), C.byref(vp), C.byref(vl))
        voff = vp.value - base
        out.append((ks, key, seg[voff:voff + vl.value]))
    return out


def spill_record_bytes(index):
    L = lib()
    arr = (C.c_int64 * len(index))(*[int(x) for x in index])
    out = Buf()
    L.tzo_buf_init(C.byref(out))
    L.tzo_spill_record_bytes(arr, len(index) // 3, C.byref(out))
    data = out.bytes()
    L.tzo_buf_free(C.byref(out))
    return data


def sorter_conf(num_partitions, cmp_kind=CMP_BYTES, partitioner=PART_HASH, send_empty=True, rle_policy=-1,
                span_records=1 << 20, sort_threads=2, legacy=False):
    return SorterConf(num_partitions, cmp_kind, partitioner, 1 if send_empty else 0, rle_policy, span_records,
                      sort_threads, 1 if legacy else 0)


def _result_dict(res, P):
    d = dict(file_out=res.file_out.bytes(), index_out=res.index_out.bytes(),
             index=np.ctypeslib.as_array(res.index, shape=(P, 3)).copy(), rle_used=bool(res.rle_used),
             eq=res.eq, total=res.total,
             counters=dict(OUTPUT_RECORDS=res.output_records, OUTPUT_BYTES=res.output_bytes,
                           OUTPUT_BYTES_WITH_OVERHEAD=res.output_bytes_with_overhead,
                           OUTPUT_BYTES_PHYSICAL=res.output_bytes_physical, SPILLED_RECORDS=res.spilled_records))
    lib().tzo_sorter_result_free(C.byref(res))
    return d


def pipelined_sort(conf, kv, key_off, key_len, val_len, partition=None):
    kv = np.ascontiguousarray(np.frombuffer(kv, dtype=np.uint8) if isinstance(kv, (bytes, bytearray)) else kv)
    key_off = np.ascontiguousarray(key_off, dtype=np.uint64)
    key_len = np.ascontiguousarray(key_len, dtype=np.uint32)
    val_len = np.ascontiguousarray(val_len, dtype=np.uint32)
    part_p = None
    if partition is not None:
        partition = np.ascontiguousarray(partition, dtype=np.int32)
        part_p = partition.ctypes.data
    res = SorterResult()
    rc = lib().tzo_pipelined_sort(C.byref(conf), kv.ctypes.data, key_off.ctypes.data, key_len.ctypes.data,
                                  val_len.ctypes.data, part_p, len(key_off), C.byref(res))
    if rc != 0:
        lib().tzo_sorter_result_free(C.byref(res))
        raise IOError("tzo_pipelined_sort rc=%d" % rc)
    return _result_dict(res, conf.num_partitions)


def unordered_write(conf, kv, key_off, key_len, val_len, partition=None):
    """UnorderedPartitionedKVWriter restatement (single buffer, no spill): per partition newest record first, no RLE,
    all-zero index entries for partitions without records.  Same result dict as pipelined_sort."""
    L = lib()
    L.tzo_unordered_write.argtypes = [C.POINTER(SorterConf), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_uint64, C.POINTER(SorterResult)]
    L.tzo_unordered_write.restype = C.c_int
    kv = np.ascontiguousarray(np.frombuffer(kv, dtype=np.uint8) if isinstance(kv, (bytes, bytearray)) else kv)
    key_off = np.ascontiguousarray(key_off, dtype=np.uint64)
    key_len = np.ascontiguousarray(key_len, dtype=np.uint32)
    val_len = np.ascontiguousarray(val_len, dtype=np.uint32)
    part_p = None
    if partition is not None:
        partition = np.ascontiguousarray(partition, dtype=np.int32)
        part_p = partition.ctypes.data
    res = SorterResult()
    rc = L.tzo_unordered_write(C.byref(conf), kv.ctypes.data if kv.size else None, key_off.ctypes.data, key_len.ctypes.data,
                               val_len.ctypes.data, part_p, len(key_off), C.byref(res))
    if rc != 0:
        L.tzo_sorter_result_free(C.byref(res))
        raise IOError("tzo_unordered_write rc=%d" % rc)
    return _result_dict(res, conf.num_partitions)


def pipelined_sort_fixed(conf, kv, klen, vlen):
    kv = np.ascontiguousarray(kv, dtype=np.uint8)
    n = kv.size // (klen + vlen)
    res = SorterResult()
    rc = lib().tzo_pipelined_sort_fixed(C.byref(conf), kv.ctypes.data, klen, vlen, n, C.byref(res))
    if rc != 0:
        lib().tzo_sorter_result_free(C.byref(res))
        raise IOError("tzo_pipelined_sort_fixed rc=%d" % rc)
    return _result_dict(res, conf.num_partitions)


def merge(segments, cmp_kind, factor=100, sort_segments=False, check_for_same_keys=True, writer_rle=False,
          has_header=True):
    """segments: list of bytes. Returns dict(records=[(key,val,same)], ifile=bytes, raw_len, comp_len, passes)."""
    L = lib()
    keep = [C.create_string_buffer(s, len(s)) for s in segments]
    arr = (Segment * max(1, len(segments)))()
    for i, s in enumerate(segments):
        arr[i].data = C.cast(keep[i], C.c_void_p)
        arr[i].len = len(s)
        arr[i].has_header = 1 if has_header else 0
    res = MergeResult()
    rc = L.tzo_merge(arr, len(segments), cmp_kind, factor, 1 if sort_segments else 0, 1 if check_for_same_keys else 0,
                     1 if writer_rle else 0, C.byref(res))
    if rc != 0:
        raise IOError("tzo_merge rc=%d" % rc)
    keys, vals = res.keys.bytes(), res.vals.bytes()
    recs = []
    ko = vo = 0
    for i in range(res.n):
        kl, vl = res.key_len[i], res.val_len[i]
        recs.append((keys[ko:ko + kl], vals[vo:vo + vl], bool(res.same_key[i])))
        ko += kl
        vo += vl
    out = dict(records=recs, ifile=res.ifile.bytes(), raw_len=res.raw_len, comp_len=res.comp_len, passes=res.passes)
    L.tzo_merge_result_free(C.byref(res))
    return out


def gen_c2(first_index, n, seed=2, threads=1):
    """SURVEY 8(d) C2 generator: n records x 80 B (16 B key + 64 B value); threads > 1 fills slices concurrently (every
    record is a function of (seed, index) alone)."""
    out = np.empty(n * 80, dtype=np.uint8)
    L = lib()
    if threads <= 1 or n < (1 << 20):
        L.tzo_gen_c2(out.ctypes.data, first_index, n, seed)
        return out
    from concurrent.futures import ThreadPoolExecutor
    per = -(-n // threads)
    base = out.ctypes.data

    def fill(t):
        a, b = t * per, min(n, (t + 1) * per)
        if b > a:
            L.tzo_gen_c2(base + a * 80, first_index + a, b - a, seed)    # ctypes releases the GIL

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(fill, range(threads)))
    return out


def bench_pipelined_fixed(conf, kv, klen, vlen, tasks):
    kv = np.ascontiguousarray(kv, dtype=np.uint8)
    n = kv.size // (klen + vlen)
    ob = C.c_uint64()
    secs = lib().tzo_bench_pipelined_fixed(C.byref(conf), kv.ctypes.data, klen, vlen, n, tasks, C.byref(ob))
    return secs, ob.value


def gen_c3_segments(nseg, seg_bytes, seed=3, threads=8, id_bits=24):
    """SURVEY 8(d) C3 generator: nseg sorted IFile segments of about seg_bytes each (Text words from a 2^24-id space,
    8-byte value = f(word)).  Returns (list of uint8 arrays, list of record counts)."""
    L = lib()
    L.tzo_gen_c3_segments.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.POINTER(Buf), C.POINTER(C.c_uint64)]
    L.tzo_gen_c3_segments.restype = None
    bufs = (Buf * nseg)()
    nrec = (C.c_uint64 * nseg)()
    L.tzo_gen_c3_segments(seed, nseg, seg_bytes, id_bits, threads, bufs, nrec)
    out = []
    for b in bufs:
        a = np.ctypeslib.as_array(b.data, shape=(b.len,)).copy()
        out.append(a)
        L.tzo_buf_free(C.byref(b))
    return out, [int(x) for x in nrec]


def merge_ifile(segments, cmp_kind, factor=100, sort_segments=False, check_for_same_keys=True, writer_rle=False):
    """TezMerger.merge + writeFile for LARGE inputs: returns (ifile bytes as a uint8 array, records, seconds); no
    per-record Python objects."""
    L = lib()
    L.tzo_merge_ifile.argtypes = [C.POINTER(Segment), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(MergeResult), C.POINTER(C.c_double)]
    L.tzo_merge_ifile.restype = C.c_int
    arr = (Segment * max(1, len(segments)))()
    keep = []
    for i, s in enumerate(segments):
        a = np.ascontiguousarray(np.frombuffer(s, dtype=np.uint8) if isinstance(s, (bytes, bytearray)) else s)
        keep.append(a)
        arr[i].data, arr[i].len, arr[i].has_header = a.ctypes.data, a.size, 1
    res = MergeResult()
    secs = C.c_double()
    rc = L.tzo_merge_ifile(arr, len(segments), cmp_kind, factor, 1 if sort_segments else 0, 1 if check_for_same_keys else 0,
                           1 if writer_rle else 0, C.byref(res), C.byref(secs))
    if rc != 0:
        raise IOError("tzo_merge rc=%d" % rc)
    out = np.ctypeslib.as_array(res.ifile.data, shape=(res.ifile.len,)).copy()
    n = int(res.n)
    L.tzo_merge_result_free(C.byref(res))
    return out, n, secs.value
