/*
 * tez_oracle.h -- CPU restatement of the Tez shuffle sort/merge hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library, and only as the checker or
 * the reported CPU baseline.  tez_b200/ never links, imports or calls it.
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference/tez-runtime-library/src/main/java/
 *  org/apache/tez/runtime/library/, abbreviated RL/; SORT/ = RL/common/sort/impl/).
 *
 * Parity pinning status (see DESIGN.md "Oracle"):
 *   - vint / record framing / EOF markers / CRC scope / rawLength: PINNED by the
 *     reference fixture TestIFile_concatenated_compressed.bin (tests/golden/).
 *   - empty segment bytes + file-length arithmetic: PINNED by
 *     TestDefaultSorter.testEmptyCaseFileLengths / TestPipelinedSorter:239-243.
 *   - merge order + SAME/DIFF key flags: PINNED by TestTezMerger literal tables.
 *   - comparator / proxy consistency: PINNED by TestProxyComparator key table.
 *   - hashBytes partition numbers, QuickSort/PriorityQueue tie order among
 *     equal keys, the needsRLE `eq` counter: PARITY UNPINNED (hadoop-common
 *     3.4.2 is not under /root/reference and no JVM exists here; restated from
 *     the published algorithm, no reference test asserts specific values).
 */
#ifndef TEZ_ORACLE_H
#define TEZ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- comparator / partitioner ids (shared numbering with include/tezgpu.h) ---- */
enum {
  TZO_CMP_BYTES = 0,         /* TezBytesComparator / raw bytes (TezBytesWritableSerialization) */
  TZO_CMP_TEXT = 1,          /* hadoop Text.Comparator: skip vint length prefix */
  TZO_CMP_BYTESWRITABLE = 2, /* hadoop BytesWritable.Comparator: skip 4-byte BE length */
  TZO_CMP_INT = 3,           /* IntWritable.Comparator: 4-byte BE signed */
  TZO_CMP_LONG = 4,          /* LongWritable.Comparator: 8-byte BE signed */
  TZO_CMP_SIGNED_BYTES = 5   /* java.nio.ByteBuffer.compareTo (signed bytes): TestTezMerger.CustomComparator */
};

enum {
  TZO_PART_GIVEN = 0,        /* caller supplies partition ids */
  TZO_PART_HASH = 1          /* HashPartitioner over the key kind's hashCode */
};

/* ---- growable byte buffer ---- */
typedef struct {
  uint8_t *data;
  size_t len, cap;
} tzo_buf;

void tzo_buf_init(tzo_buf *b);
void tzo_buf_free(tzo_buf *b);
void tzo_buf_put(tzo_buf *b, const void *p, size_t n);

/* ---- hadoop WritableUtils vint/vlong (Appendix A.2 of SURVEY; call sites SORT/IFile.java:398-399,561,575-576) ---- */
int tzo_vint_size(int64_t v);
int tzo_write_vlong(uint8_t *dst, int64_t v);            /* returns bytes written (1..9) */
int tzo_read_vlong(const uint8_t *src, int64_t *out);    /* returns bytes consumed */
int tzo_decode_vint_size(uint8_t first);

/* ---- CRC-32 (DataChecksum.Type.CRC32 / PureJavaCrc32 == zlib crc32), SORT/IFileOutputStream.java:53-90 ---- */
uint32_t tzo_crc32(uint32_t crc, const uint8_t *p, size_t n);

/* ---- WritableComparator.hashBytes + HashPartitioner (RL/partitioner/HashPartitioner.java:33-35) ---- */
int32_t tzo_hash_bytes(const uint8_t *p, size_t n);
int32_t tzo_key_hash(int cmp_kind, const uint8_t *key, size_t klen);
int32_t tzo_hash_partition(int32_t hash, int32_t num_partitions);

/* ---- RawComparators (RL/utils/FastByteComparisons.java:92-116, RL/common/comparator/TezBytesComparator.java:37-41) ---- */
int tzo_compare(int cmp_kind, const uint8_t *a, int la, const uint8_t *b, int lb);
/* TezBytesComparator.getProxy (RL/common/comparator/TezBytesComparator.java:43-61) */
int32_t tzo_bytes_proxy(const uint8_t *content, int len);
/* PipelinedSorter prefix word (SORT/PipelinedSorter.java:164,316-323,450-456) */
int32_t tzo_pipelined_prefix(int32_t partition, int32_t proxy, int32_t num_partitions);

/* ---- IFile.Writer (SORT/IFile.java:262-634) ---- */
typedef struct {
  tzo_buf *out;          /* rawOut */
  size_t start;          /* rawOut.getPos() at construction */
  int rle;
  int prev_is_repeat;    /* prevKey == REPEAT_KEY */
  tzo_buf previous;      /* previous key (only kept when rle) */
  uint32_t crc;          /* IFileOutputStream running checksum over the body */
  int64_t raw_len;       /* decompressedBytesWritten */
  int64_t comp_len;      /* compressedBytesWritten */
  int64_t records;
  int64_t rle_written;
  int closed;
} tzo_ifile_writer;

void tzo_writer_open(tzo_ifile_writer *w, tzo_buf *out, int rle);
/* key==NULL means IFile.REPEAT_KEY (append(DataInputBuffer,DataInputBuffer), :534-557) */
void tzo_writer_append(tzo_ifile_writer *w, const uint8_t *key, int klen, const uint8_t *val, int vlen);
void tzo_writer_close(tzo_ifile_writer *w);

/* ---- IFile.Reader record cursor (SORT/IFile.java:877-1000); in-memory body per OG/InMemoryReader.java:142-254 ---- */
enum { TZO_NO_KEY = 0, TZO_NEW_KEY = 1, TZO_SAME_KEY = 2 };
typedef struct {
  const uint8_t *data;   /* segment bytes */
  size_t pos, end;       /* body cursor [pos,end): end excludes the 4-byte checksum */
  int64_t cur_klen, cur_vlen, prev_klen, orig_klen;
  const uint8_t *key;    /* current key bytes (points into data) */
  int klen;
  int eof;
  int64_t records;
  size_t length;         /* segment length for sort-by-size (getLength) */
} tzo_ifile_reader;

/* has_header: 1 for on-disk segment (TIF\0 + body + crc), 0 for in-memory (body + crc). verify_crc: check trailer.
 * returns 0 ok, <0 on malformed / checksum mismatch */
int tzo_reader_open(tzo_ifile_reader *r, const uint8_t *seg, size_t len, int has_header, int verify_crc);
int tzo_reader_read_raw_key(tzo_ifile_reader *r);                                 /* KeyState */
void tzo_reader_next_raw_value(tzo_ifile_reader *r, const uint8_t **val, int *vlen);

/* ---- TezSpillRecord (SORT/TezSpillRecord.java:48-52,111-146) ---- */
/* idx: P triples (start, rawLen, partLen); writes P*24+8 bytes big-endian + CRC long */
void tzo_spill_record_bytes(const int64_t *idx, int num_partitions, tzo_buf *out);

/* ---- sorter restatements ---- */
typedef struct {
  int num_partitions;
  int cmp_kind;
  int partitioner;            /* TZO_PART_* */
  int send_empty_partition_details; /* tez.runtime.empty.partitions.info-via-events.enabled (default 1) */
  int rle_policy;             /* -1 = reference rule (eq>0.1*total from QuickSort compare counts), 0 = off, 1 = on */
  int span_records;           /* PipelinedSorter: records per sort span (reference sizes spans for <=1M records, :229,346-360) */
  int sort_threads;           /* span sorts run on this many threads (tez.runtime.pipelined.sorter.sort.threads, default 2) */
  int legacy;                 /* 1 = DefaultSorter semantics (dflt/DefaultSorter.java) */
} tzo_sorter_conf;

typedef struct {
  tzo_buf file_out;           /* file.out bytes */
  tzo_buf index_out;          /* file.out.index bytes */
  int64_t *index;             /* P triples */
  int rle_used;
  int64_t eq, total;
  /* counters (SORT/ExternalSorter.java:141-167) */
  int64_t output_records, output_bytes, output_bytes_with_overhead, output_bytes_physical, spilled_records;
} tzo_sorter_result;

/* records: kv buffer + per-record key offset/len and value len (value follows key); partition[] used when partitioner==GIVEN */
int tzo_pipelined_sort(const tzo_sorter_conf *conf, const uint8_t *kv, const uint64_t *key_off,
                       const uint32_t *key_len, const uint32_t *val_len, const int32_t *partition,
                       uint64_t n, tzo_sorter_result *res);
void tzo_sorter_result_free(tzo_sorter_result *res);

/* UnorderedPartitionedKVWriter (RL/common/writers/UnorderedPartitionedKVWriter.java:459-472,688-703,1058-1144), single
 * buffer / no spill: per partition the records newest first, IFile without run-length encoding, all-zero index entries
 * for partitions without records */
int tzo_unordered_write(const tzo_sorter_conf *conf, const uint8_t *kv, const uint64_t *key_off, const uint32_t *key_len,
                        const uint32_t *val_len, const int32_t *partition, uint64_t n, tzo_sorter_result *res);

/* fixed-width convenience used by bench cpu_baseline: n records of (klen+vlen) bytes packed back to back */
int tzo_pipelined_sort_fixed(const tzo_sorter_conf *conf, const uint8_t *kv, uint32_t klen, uint32_t vlen,
                             uint64_t n, tzo_sorter_result *res);

/* ---- TezMerger.MergeQueue (SORT/TezMerger.java:465-1065) ---- */
typedef struct {
  const uint8_t *data;
  size_t len;
  int has_header;             /* DiskSegment (1) vs in-memory body+crc (0) */
} tzo_segment;

typedef struct {
  /* merged stream */
  uint64_t n;
  tzo_buf keys, vals;         /* concatenated bytes */
  uint32_t *key_len, *val_len;
  uint8_t *same_key;          /* isSameKey() per emitted record */
  /* writeFile output (IFile segment incl. header+crc) */
  tzo_buf ifile;
  int64_t raw_len, comp_len;
  int passes;                 /* number of intermediate merge passes materialised */
} tzo_merge_result;

/* restates TezMerger.merge(...) + iteration (+ writeFile into an IFile.Writer(rle=writer_rle)) */
int tzo_merge(const tzo_segment *segs, int nseg, int cmp_kind, int factor, int sort_segments,
              int check_for_same_keys, int writer_rle, tzo_merge_result *res);
void tzo_merge_result_free(tzo_merge_result *res);

/* ---- synthetic generators shared by oracle, tests and bench (splitmix64 counter based) ---- */
uint64_t tzo_splitmix64(uint64_t x);
/* C2 generator: record i = 16B key (two BE splitmix64 words) + 64B value (8 BE words) */
void tzo_gen_c2(uint8_t *dst, uint64_t first_index, uint64_t n, uint64_t seed);

/* same, producing only the writeFile output + record count (res->n) -- no per-record stream; wall time in *seconds */
int tzo_merge_ifile(const tzo_segment *segs, int nseg, int cmp_kind, int factor, int sort_segments,
                    int check_for_same_keys, int writer_rle, tzo_merge_result *res, double *seconds);

/* C3 generator: sorted IFile segments of (Text word from a 2^24-id space, 8-byte value = f(word)) */
int tzo_c3_word(uint32_t id, uint8_t *out);
void tzo_gen_c3_segment(uint64_t seed, uint32_t seg_index, uint64_t target_bytes, int id_bits, tzo_buf *out, uint64_t *nrecords);
void tzo_gen_c3_segments(uint64_t seed, uint32_t nseg, uint64_t target_bytes, int id_bits, int threads, tzo_buf *outs, uint64_t *nrec);

/* multi-threaded CPU baseline: T independent PipelinedSorter tasks over slices (one task per thread, as Tez runs one map task per core) */
double tzo_bench_pipelined_fixed(const tzo_sorter_conf *conf, const uint8_t *kv, uint32_t klen, uint32_t vlen,
                                 uint64_t n, int tasks, uint64_t *out_bytes);

#ifdef __cplusplus
}
#endif
#endif
