/*
 * tez_oracle.c -- CPU restatement of the Tez shuffle sort/merge hot path (plain C).
 *
 * TEST INFRASTRUCTURE ONLY -- see tez_oracle.h.  Not linked into, imported by
 * or called from the product (tez_b200/).  Used by tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference leg as the checker / baseline.
 *
 * Citations: RL/ = /root/reference/tez-runtime-library/src/main/java/org/apache/
 * tez/runtime/library/, SORT/ = RL/common/sort/impl/.  hadoop-common 3.4.2
 * classes (WritableUtils, QuickSort, HeapSort, util.PriorityQueue,
 * WritableComparator) are NOT under /root/reference (pom.xml:79); they are
 * restated from their published algorithms -- see SURVEY.md Appendix A.2.
 */
#include "tez_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ buffers */
void tzo_buf_init(tzo_buf *b) { b->data = NULL; b->len = 0; b->cap = 0; }
void tzo_buf_free(tzo_buf *b) { free(b->data); b->data = NULL; b->len = b->cap = 0; }
static void tzo_buf_reserve(tzo_buf *b, size_t extra) {
  if (b->len + extra <= b->cap) return;
  size_t nc = b->cap ? b->cap : 256;
  while (nc < b->len + extra) nc = nc + nc / 2 + 64;
  b->data = (uint8_t *)realloc(b->data, nc);
  if (!b->data) { fprintf(stderr, "tez_oracle: out of memory\n"); abort(); }
  b->cap = nc;
}
void tzo_buf_put(tzo_buf *b, const void *p, size_t n) {
  tzo_buf_reserve(b, n);
  if (n) memcpy(b->data + b->len, p, n);
  b->len += n;
}

/* ------------------------------------------------------------------ vint
 * hadoop WritableUtils.writeVLong / readVLong / decodeVIntSize / getVIntSize. */
int tzo_vint_size(int64_t i) {
  if (i >= -112 && i <= 127) return 1;
  if (i < 0) i = ~i;
  int data_bits = 64 - __builtin_clzll((unsigned long long)i | 1ULL);
  if (i == 0) data_bits = 0;
  return (data_bits + 7) / 8 + 1;
}
int tzo_write_vlong(uint8_t *dst, int64_t i) {
  if (i >= -112 && i <= 127) { dst[0] = (uint8_t)(int8_t)i; return 1; }
  int len = -112;
  if (i < 0) { i = ~i; len = -120; }
  int64_t tmp = i;
  while (tmp != 0) { tmp = (int64_t)((uint64_t)tmp >> 8); len--; }
  dst[0] = (uint8_t)(int8_t)len;
  len = (len < -120) ? -(len + 120) : -(len + 112);
  for (int idx = len; idx != 0; idx--) {
    int shift = (idx - 1) * 8;
    dst[1 + (len - idx)] = (uint8_t)(((uint64_t)i >> shift) & 0xFF);
  }
  return len + 1;
}
int tzo_decode_vint_size(uint8_t first) {
  int8_t v = (int8_t)first;
  if (v >= -112) return 1;
  if (v < -120) return -119 - v;
  return -111 - v;
}
int tzo_read_vlong(const uint8_t *src, int64_t *out) {
  int8_t first = (int8_t)src[0];
  int len = tzo_decode_vint_size(src[0]);
  if (len == 1) { *out = first; return 1; }
  uint64_t i = 0;
  for (int idx = 0; idx < len - 1; idx++) i = (i << 8) | src[1 + idx];
  int negative = (first < -120) || (first >= -112 && first < 0);
  *out = negative ? (int64_t)~i : (int64_t)i;
  return len;
}

/* ------------------------------------------------------------------ crc32 (slice-by-8; poly 0xEDB88320) */
static uint32_t crc_tab[8][256];
static pthread_once_t crc_once = PTHREAD_ONCE_INIT;
static void crc_init(void) {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : (c >> 1);
    crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; i++)
    for (int t = 1; t < 8; t++) crc_tab[t][i] = (crc_tab[t - 1][i] >> 8) ^ crc_tab[0][crc_tab[t - 1][i] & 0xFF];
}
uint32_t tzo_crc32(uint32_t crc, const uint8_t *p, size_t n) {
  pthread_once(&crc_once, crc_init);
  uint32_t c = ~crc;
  while (n && ((uintptr_t)p & 7)) { c = crc_tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8); n--; }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = crc_tab[7][w & 0xFF] ^ crc_tab[6][(w >> 8) & 0xFF] ^ crc_tab[5][(w >> 16) & 0xFF] ^ crc_tab[4][(w >> 24) & 0xFF] ^
        crc_tab[3][(w >> 32) & 0xFF] ^ crc_tab[2][(w >> 40) & 0xFF] ^ crc_tab[1][(w >> 48) & 0xFF] ^ crc_tab[0][(w >> 56) & 0xFF];
    p += 8; n -= 8;
  }
  while (n--) c = crc_tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return ~c;
}

/* ------------------------------------------------------------------ hash / partition */
int32_t tzo_hash_bytes(const uint8_t *p, size_t n) {
  /* WritableComparator.hashBytes: h = 1; h = 31*h + (signed byte) */
  uint32_t h = 1;
  for (size_t i = 0; i < n; i++) h = 31u * h + (uint32_t)(int32_t)(int8_t)p[i];
  return (int32_t)h;
}
static int key_skip(int cmp_kind, const uint8_t *key, size_t klen) {
  if (klen == 0) return 0;
  if (cmp_kind == TZO_CMP_TEXT) return tzo_decode_vint_size(key[0]);
  if (cmp_kind == TZO_CMP_BYTESWRITABLE) return 4;
  return 0;
}
int32_t tzo_key_hash(int cmp_kind, const uint8_t *key, size_t klen) {
  if (cmp_kind == TZO_CMP_INT && klen >= 4)
    return (int32_t)(((uint32_t)key[0] << 24) | ((uint32_t)key[1] << 16) | ((uint32_t)key[2] << 8) | key[3]);
  if (cmp_kind == TZO_CMP_LONG && klen >= 8) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | key[i];
    return (int32_t)(uint32_t)(v ^ (v >> 32));   /* LongWritable.hashCode */
  }
  int s = key_skip(cmp_kind, key, klen);
  if ((size_t)s > klen) s = (int)klen;
  return tzo_hash_bytes(key + s, klen - (size_t)s);
}
int32_t tzo_hash_partition(int32_t hash, int32_t P) {
  /* RL/partitioner/HashPartitioner.java:33-35 */
  return (hash & 0x7fffffff) % P;
}

/* ------------------------------------------------------------------ comparators */
static int compare_bytes(const uint8_t *a, int la, const uint8_t *b, int lb) {
  /* RL/utils/FastByteComparisons.java:92-116 (PureJavaComparer): unsigned lexicographic, then l1-l2 */
  int n = la < lb ? la : lb;
  for (int i = 0; i < n; i++) {
    int x = a[i], y = b[i];
    if (x != y) return x - y;
  }
  return la - lb;
}
int tzo_compare(int kind, const uint8_t *a, int la, const uint8_t *b, int lb) {
  switch (kind) {
    case TZO_CMP_TEXT: {
      int n1 = la ? tzo_decode_vint_size(a[0]) : 0, n2 = lb ? tzo_decode_vint_size(b[0]) : 0;
      return compare_bytes(a + n1, la - n1, b + n2, lb - n2);
    }
    case TZO_CMP_BYTESWRITABLE:
      return compare_bytes(a + 4, la - 4, b + 4, lb - 4);
    case TZO_CMP_INT: {
      int32_t x = (int32_t)(((uint32_t)a[0] << 24) | ((uint32_t)a[1] << 16) | ((uint32_t)a[2] << 8) | a[3]);
      int32_t y = (int32_t)(((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]);
      return x < y ? -1 : (x == y ? 0 : 1);
    }
    case TZO_CMP_LONG: {
      uint64_t ux = 0, uy = 0;
      for (int i = 0; i < 8; i++) { ux = (ux << 8) | a[i]; uy = (uy << 8) | b[i]; }
      int64_t x = (int64_t)ux, y = (int64_t)uy;
      return x < y ? -1 : (x == y ? 0 : 1);
    }
    case TZO_CMP_SIGNED_BYTES: {
      /* java.nio.ByteBuffer.compareTo as used by TestTezMerger.CustomComparator */
      int n = la < lb ? la : lb;
      for (int i = 0; i < n; i++) {
        int x = (int8_t)a[i], y = (int8_t)b[i];
        if (x != y) return x < y ? -1 : 1;
      }
      return la - lb;
    }
    default:
      return compare_bytes(a, la, b, lb);
  }
}
int32_t tzo_bytes_proxy(const uint8_t *c, int len) {
  /* RL/common/comparator/TezBytesComparator.java:43-61 */
  int b1 = 0, b2 = 0, b3 = 0;
  if (len >= 3) b3 = c[2];
  if (len >= 2) b2 = c[1];
  if (len >= 1) b1 = c[0];
  return (b1 << 16) | (b2 << 8) | b3;
}
static int bitcount(int n) { int bit = 0; while (n != 0) { bit++; n >>= 1; } return bit; }
int32_t tzo_pipelined_prefix(int32_t partition, int32_t proxy, int32_t P) {
  /* SORT/PipelinedSorter.java:164 (partitionBits = bitcount(partitions)+1), :456 */
  int partition_bits = bitcount(P) + 1;
  return (int32_t)(((uint32_t)partition << (32 - partition_bits)) | ((uint32_t)proxy >> partition_bits));
}

/* ------------------------------------------------------------------ IFile.Writer */
static const uint8_t IFILE_HEADER[4] = {'T', 'I', 'F', 0};

static void w_body(tzo_ifile_writer *w, const uint8_t *p, size_t n) {
  /* bytes that pass through IFileOutputStream: checksummed (SORT/IFileOutputStream.java:81-90) */
  w->crc = tzo_crc32(w->crc, p, n);
  tzo_buf_put(w->out, p, n);
}
static void w_vint(tzo_ifile_writer *w, int64_t v) {
  uint8_t tmp[10];
  int n = tzo_write_vlong(tmp, v);
  w_body(w, tmp, (size_t)n);
}
void tzo_writer_open(tzo_ifile_writer *w, tzo_buf *out, int rle) {
  memset(w, 0, sizeof(*w));
  w->out = out;
  w->start = out->len;
  w->rle = rle;
  tzo_buf_init(&w->previous);
  /* writeHeader: header bypasses the checksum stream (SORT/IFile.java:373-379) */
  tzo_buf_put(out, IFILE_HEADER, 4);
}
static void w_value_marker(tzo_ifile_writer *w) {
  /* writeValueMarker :604-614 */
  if (w->prev_is_repeat) { w_vint(w, -3); w->raw_len += 1; }
}
void tzo_writer_append(tzo_ifile_writer *w, const uint8_t *key, int klen, const uint8_t *val, int vlen) {
  /* append(DataInputBuffer key, DataInputBuffer value) :534-557 */
  int same = (key == NULL);
  if (!same && w->rle) {
    same = (klen != 0) && (compare_bytes(w->previous.data, (int)w->previous.len, key, klen) == 0);
  }
  if (!same) {
    /* writeKVPair :572-587 */
    w_value_marker(w);
    w_vint(w, klen);
    w_vint(w, vlen);
    w_body(w, key, (size_t)klen);
    w_body(w, val, (size_t)vlen);
    w->raw_len += klen + vlen + tzo_vint_size(klen) + tzo_vint_size(vlen);
    if (w->rle) { w->previous.len = 0; tzo_buf_put(&w->previous, key, (size_t)klen); }
  } else {
    /* writeValue :559-570 + writeRLE :589-602 */
    if (!w->prev_is_repeat) { w_vint(w, -2); w->raw_len += 1; w->rle_written++; }
    w_vint(w, vlen);
    w_body(w, val, (size_t)vlen);
    w->raw_len += vlen + tzo_vint_size(vlen);
  }
  w->prev_is_repeat = same;
  w->records++;
}
void tzo_writer_close(tzo_ifile_writer *w) {
  /* close() :381-435 */
  if (w->closed) return;
  w->closed = 1;
  w_value_marker(w);
  w_vint(w, -1);
  w_vint(w, -1);
  w->raw_len += 2;
  w->raw_len += 4; /* header */
  uint8_t c[4] = {(uint8_t)(w->crc >> 24), (uint8_t)(w->crc >> 16), (uint8_t)(w->crc >> 8), (uint8_t)w->crc};
  tzo_buf_put(w->out, c, 4); /* checksumOut.finish(): big-endian CRC32 of the body */
  w->comp_len = (int64_t)(w->out->len - w->start);
  tzo_buf_free(&w->previous);
}

/* ------------------------------------------------------------------ IFile.Reader */
int tzo_reader_open(tzo_ifile_reader *r, const uint8_t *seg, size_t len, int has_header, int verify_crc) {
  memset(r, 0, sizeof(*r));
  r->data = seg;
  r->length = len;
  size_t body0 = 0;
  if (has_header) {
    if (len < 4 || seg[0] != 'T' || seg[1] != 'I' || seg[2] != 'F') return -1; /* verifyHeaderMagic :1004-1009 */
    if (seg[3] != 0) return -2; /* compressed segments are out of scope for the oracle */
    body0 = 4;
  }
  if (len < body0 + 4) return -3;
  r->pos = body0;
  r->end = len - 4;
  if (verify_crc) {
    /* SORT/IFileInputStream.java:235-289: trailer == CRC32(body) */
    uint32_t c = tzo_crc32(0, seg + body0, r->end - body0);
    uint32_t t = ((uint32_t)seg[len - 4] << 24) | ((uint32_t)seg[len - 3] << 16) | ((uint32_t)seg[len - 2] << 8) | seg[len - 1];
    if (c != t) return -4;
  }
  return 0;
}
static int r_vint(tzo_ifile_reader *r, int64_t *v) {
  if (r->pos >= r->end) return -1;
  int n = tzo_decode_vint_size(r->data[r->pos]);
  if (r->pos + (size_t)n > r->end) return -1;
  tzo_read_vlong(r->data + r->pos, v);
  r->pos += (size_t)n;
  return 0;
}
int tzo_reader_read_raw_key(tzo_ifile_reader *r) {
  /* positionToNextRecord :903-935 + readRawKey :954-978 */
  if (r->eof) return TZO_NO_KEY;
  r->prev_klen = r->cur_klen;
  if (r->prev_klen == -2) {
    /* readValueLength :877-883 */
    if (r_vint(r, &r->cur_vlen)) { r->eof = 1; return -1; }
    if (r->cur_vlen == -3) {
      if (r_vint(r, &r->cur_klen) || r_vint(r, &r->cur_vlen)) { r->eof = 1; return -1; }
      if (r->cur_klen != -2) r->orig_klen = r->cur_klen;
    }
  } else {
    /* readKeyValueLength :885-894 */
    if (r_vint(r, &r->cur_klen) || r_vint(r, &r->cur_vlen)) { r->eof = 1; return -1; }
    if (r->cur_klen != -2) r->orig_klen = r->cur_klen;
  }
  if (r->cur_klen == -1 && r->cur_vlen == -1) { r->eof = 1; return TZO_NO_KEY; }
  if (r->cur_klen != -2 && r->cur_klen < 0) { r->eof = 1; return -1; }
  if (r->cur_vlen < 0) { r->eof = 1; return -1; }
  if (r->cur_klen == -2) {
    r->klen = (int)r->orig_klen;  /* key.reset(keyBytes, originalKeyLength) */
    return TZO_SAME_KEY;
  }
  if (r->pos + (size_t)r->cur_klen > r->end) { r->eof = 1; return -1; }
  r->key = r->data + r->pos;
  r->klen = (int)r->cur_klen;
  r->pos += (size_t)r->cur_klen;
  return TZO_NEW_KEY;
}
void tzo_reader_next_raw_value(tzo_ifile_reader *r, const uint8_t **val, int *vlen) {
  /* nextRawValue :980-1002 */
  *val = r->data + r->pos;
  *vlen = (int)r->cur_vlen;
  r->pos += (size_t)r->cur_vlen;
  r->records++;
}

/* ------------------------------------------------------------------ TezSpillRecord */
static void put_be64(uint8_t *d, uint64_t v) { for (int i = 0; i < 8; i++) d[i] = (uint8_t)(v >> (56 - 8 * i)); }
void tzo_spill_record_bytes(const int64_t *idx, int P, tzo_buf *out) {
  /* SORT/TezSpillRecord.java:111-146: LongBuffer view of a big-endian ByteBuffer, then writeLong(crc) */
  size_t n = (size_t)P * 24;
  uint8_t *tmp = (uint8_t *)malloc(n + 8);
  for (int p = 0; p < P; p++)
    for (int k = 0; k < 3; k++) put_be64(tmp + (size_t)p * 24 + (size_t)k * 8, (uint64_t)idx[p * 3 + k]);
  uint32_t c = tzo_crc32(0, tmp, n);
  put_be64(tmp + n, (uint64_t)c);
  tzo_buf_put(out, tmp, n + 8);
  free(tmp);
}

/* ------------------------------------------------------------------ hadoop util.QuickSort / HeapSort over an IndexedSortable
 * (restated; call sites SORT/ExternalSorter.java:197-199, SORT/PipelinedSorter.java:965-973) */
typedef struct {
  int32_t prefix;     /* PipelinedSorter: partition|proxy word; DefaultSorter: partition */
  uint32_t klen, vlen;
  uint64_t koff;
} meta_t;

typedef struct {
  meta_t *m;
  const uint8_t *kv;
  int cmp_kind;
  int legacy;
  int64_t eq;
} sortable_t;

static inline int s_compare(sortable_t *s, int i, int j) {
  const meta_t *a = &s->m[i], *b = &s->m[j];
  /* SortSpan.compare :1013-1023 / DefaultSorter.compare :451-471 */
  if (a->prefix != b->prefix) return a->prefix - b->prefix;
  if (!s->legacy && (a->klen == 0 || b->klen == 0)) {
    /* compareKeys :996-1003 */
    if (a->klen == b->klen) s->eq++;
    return (int)a->klen - (int)b->klen;
  }
  int c = tzo_compare(s->cmp_kind, s->kv + a->koff, (int)a->klen, s->kv + b->koff, (int)b->klen);
  if (c == 0) s->eq++;
  return c;
}
static inline void s_swap(sortable_t *s, int i, int j) { meta_t t = s->m[i]; s->m[i] = s->m[j]; s->m[j] = t; }
static inline void s_fix(sortable_t *s, int p, int r) { if (s_compare(s, p, r) > 0) s_swap(s, p, r); }

static void heap_down(sortable_t *s, int b, int i, int N) {
  for (int idx = i << 1; idx < N; idx = i << 1) {
    if (idx + 1 < N && s_compare(s, b + idx, b + idx + 1) < 0) {
      if (s_compare(s, b + i, b + idx + 1) < 0) s_swap(s, b + i, b + idx + 1);
      else return;
      i = idx + 1;
    } else if (s_compare(s, b + i, b + idx) < 0) {
      s_swap(s, b + i, b + idx);
      i = idx;
    } else return;
  }
}
static void heap_sort(sortable_t *s, int p, int r) {
  int N = r - p;
  int t = 1;
  while ((t << 1) <= N) t <<= 1; /* Integer.highestOneBit(N) */
  for (int i = t; i > 1; i >>= 1)
    for (int j = i >> 1; j < i; ++j) heap_down(s, p - 1, j, N + 1);
  for (int i = r - 1; i > p; --i) {
    s_swap(s, p, i);
    heap_down(s, p - 1, 1, i - p + 1);
  }
}
static void quick_sort_internal(sortable_t *s, int p, int r, int depth) {
  for (;;) {
    if (r - p < 13) {
      for (int i = p; i < r; ++i)
        for (int j = i; j > p && s_compare(s, j - 1, j) > 0; --j) s_swap(s, j, j - 1);
      return;
    }
    if (--depth < 0) { heap_sort(s, p, r); return; }
    s_fix(s, (int)(((unsigned)p + (unsigned)r) >> 1), p);
    s_fix(s, (int)(((unsigned)p + (unsigned)r) >> 1), r - 1);
    s_fix(s, p, r - 1);
    int i = p, j = r, ll = p, rr = r, cr;
    for (;;) {
      while (++i < j) {
        if ((cr = s_compare(s, i, p)) > 0) break;
        if (0 == cr && ++ll != i) s_swap(s, ll, i);
      }
      while (--j > i) {
        if ((cr = s_compare(s, p, j)) > 0) break;
        if (0 == cr && --rr != j) s_swap(s, rr, j);
      }
      if (i < j) s_swap(s, i, j);
      else break;
    }
    j = i;
    while (ll >= p) s_swap(s, ll--, --i);
    while (rr < r) s_swap(s, rr++, j++);
    if (i - p < r - j) { quick_sort_internal(s, p, i, depth); p = j; }
    else { quick_sort_internal(s, j, r, depth); r = i; }
  }
}
static void hadoop_quick_sort(sortable_t *s, int p, int r) {
  if (r - p <= 1) return;
  /* QuickSort.getMaxDepth: (32 - numberOfLeadingZeros(x - 1)) << 2, x = r - p >= 2 here */
  int max_depth = (32 - __builtin_clz((unsigned)(r - p - 1))) << 2;
  quick_sort_internal(s, p, r, max_depth);
}

/* ------------------------------------------------------------------ PipelinedSorter / DefaultSorter restatement */
typedef struct {
  sortable_t s;
  int lo, hi;
} span_job;

typedef struct {
  span_job *jobs;
  int njobs;
  int next;
  pthread_mutex_t mu;
} span_pool;

static void *span_worker(void *arg) {
  span_pool *pool = (span_pool *)arg;
  for (;;) {
    pthread_mutex_lock(&pool->mu);
    int j = pool->next++;
    pthread_mutex_unlock(&pool->mu);
    if (j >= pool->njobs) return NULL;
    span_job *job = &pool->jobs[j];
    hadoop_quick_sort(&job->s, job->lo, job->hi); /* SortSpan.sort :965-973 (only when length()>1) */
  }
}

/* binary min-heap of span cursors (SpanMerger/SpanHeap :1338-1351, java.util.PriorityQueue; tie order UNPINNED:
 * we break ties by span index) */
typedef struct { int span; int pos, end; } span_cur;

static int span_less(const meta_t *m, const uint8_t *kv, int cmp_kind, const span_cur *a, const span_cur *b) {
  const meta_t *x = &m[a->pos], *y = &m[b->pos];
  int c;
  if (x->prefix != y->prefix) c = x->prefix - y->prefix;
  else c = tzo_compare(cmp_kind, kv + x->koff, (int)x->klen, kv + y->koff, (int)y->klen);
  if (c != 0) return c < 0;
  return a->span < b->span;
}

int tzo_pipelined_sort(const tzo_sorter_conf *conf, const uint8_t *kv, const uint64_t *key_off, const uint32_t *key_len,
                       const uint32_t *val_len, const int32_t *partition, uint64_t n, tzo_sorter_result *res) {
  const int P = conf->num_partitions;
  memset(res, 0, sizeof(*res));
  tzo_buf_init(&res->file_out);
  tzo_buf_init(&res->index_out);
  res->index = (int64_t *)calloc((size_t)P * 3, sizeof(int64_t));
  if (n > 0x7fffffffULL) return -1;

  /* collect :398-466 */
  meta_t *m = (meta_t *)malloc((size_t)(n ? n : 1) * sizeof(meta_t));
  int partition_bits = bitcount(P) + 1;
  for (uint64_t i = 0; i < n; i++) {
    const uint8_t *k = kv + key_off[i];
    int32_t p;
    if (conf->partitioner == TZO_PART_HASH) p = tzo_hash_partition(tzo_key_hash(conf->cmp_kind, k, key_len[i]), P);
    else p = partition[i];
    if (p < 0 || p >= P) { free(m); return -2; } /* "Illegal partition" :410-413 */
    int32_t prefix;
    if (conf->legacy) prefix = p;
    else {
      int32_t proxy = (conf->cmp_kind == TZO_CMP_BYTES) ? tzo_bytes_proxy(k, (int)key_len[i]) : 0; /* hasher only for ProxyComparator :450-454 */
      prefix = (int32_t)(((uint32_t)p << (32 - partition_bits)) | ((uint32_t)proxy >> partition_bits));
    }
    m[i].prefix = prefix;
    m[i].koff = key_off[i];
    m[i].klen = key_len[i];
    m[i].vlen = val_len[i];
    res->output_records++;
    res->output_bytes += (int64_t)key_len[i] + val_len[i];
  }

  /* sort(): spans sorted independently (thread pool :231-239,362-367), then merged by SpanMerger */
  int span_records = conf->legacy ? (int)(n ? n : 1) : (conf->span_records > 0 ? conf->span_records : (1 << 20));
  int nspans = (int)((n + (uint64_t)span_records - 1) / (uint64_t)span_records);
  if (nspans == 0) nspans = 1;
  span_pool pool;
  pool.jobs = (span_job *)calloc((size_t)nspans, sizeof(span_job));
  pool.njobs = nspans;
  pool.next = 0;
  pthread_mutex_init(&pool.mu, NULL);
  for (int s = 0; s < nspans; s++) {
    span_job *job = &pool.jobs[s];
    job->s.m = m; job->s.kv = kv; job->s.cmp_kind = conf->cmp_kind; job->s.legacy = conf->legacy; job->s.eq = 0;
    job->lo = (int)((uint64_t)s * (uint64_t)span_records);
    uint64_t hi = (uint64_t)(s + 1) * (uint64_t)span_records;
    job->hi = (int)(hi < n ? hi : n);
  }
  int nthreads = conf->sort_threads > 0 ? conf->sort_threads : 1;
  if (nthreads > nspans) nthreads = nspans;
  if (nthreads <= 1) span_worker(&pool);
  else {
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, span_worker, &pool);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
  }
  pthread_mutex_destroy(&pool.mu);
  int64_t eq = 0;
  for (int s = 0; s < nspans; s++) eq += pool.jobs[s].s.eq;
  res->eq = eq;
  res->total = (int64_t)n;

  /* needsRLE :1436-1438 (eq > 0.1*total); DefaultSorter: isRLENeeded(sameKey,totalKeys) evaluated with the
   * counters captured BEFORE the sort of the (first) spill => sameKey==0 => false (dflt/DefaultSorter.java:722-729,856-858) */
  int rle;
  if (conf->rle_policy >= 0) rle = conf->rle_policy;
  else if (conf->legacy) rle = 0;
  else rle = ((double)eq > 0.1 * (double)n);
  res->rle_used = rle;

  /* SpanMerger: k-way merge of the sorted spans into one order array */
  meta_t *sorted = m;
  if (nspans > 1) {
    sorted = (meta_t *)malloc((size_t)n * sizeof(meta_t));
    span_cur *heap = (span_cur *)malloc(sizeof(span_cur) * (size_t)nspans);
    int hs = 0;
    for (int s = 0; s < nspans; s++) {
      if (pool.jobs[s].lo >= pool.jobs[s].hi) continue;
      span_cur c = {s, pool.jobs[s].lo, pool.jobs[s].hi};
      int i = hs++;
      heap[i] = c;
      while (i > 0) {
        int par = (i - 1) / 2;
        if (!span_less(m, kv, conf->cmp_kind, &heap[i], &heap[par])) break;
        span_cur t = heap[i]; heap[i] = heap[par]; heap[par] = t; i = par;
      }
    }
    uint64_t o = 0;
    while (hs > 0) {
      sorted[o++] = m[heap[0].pos];
      heap[0].pos++;
      if (heap[0].pos >= heap[0].end) heap[0] = heap[--hs];
      int i = 0;
      for (;;) {
        int l = 2 * i + 1, r = l + 1, b = i;
        if (l < hs && span_less(m, kv, conf->cmp_kind, &heap[l], &heap[b])) b = l;
        if (r < hs && span_less(m, kv, conf->cmp_kind, &heap[r], &heap[b])) b = r;
        if (b == i) break;
        span_cur t = heap[i]; heap[i] = heap[b]; heap[b] = t; i = b;
      }
    }
    free(heap);
  }
  free(pool.jobs);

  /* spill :558-647 -- one IFile segment per partition, in partition order */
  uint64_t pos = 0;
  for (int p = 0; p < P; p++) {
    int64_t seg_start = (int64_t)res->file_out.len;
    uint64_t first = pos;
    while (pos < n) {
      int32_t rp = conf->legacy ? sorted[pos].prefix : (int32_t)((uint32_t)sorted[pos].prefix >> (32 - partition_bits));
      if (rp != p) break;
      pos++;
    }
    int has = pos > first;
    int64_t raw = 0, part = 0;
    if (has || !conf->send_empty_partition_details) {
      tzo_ifile_writer w;
      tzo_writer_open(&w, &res->file_out, rle);
      for (uint64_t i = first; i < pos; i++)
        tzo_writer_append(&w, kv + sorted[i].koff, (int)sorted[i].klen, kv + sorted[i].koff + sorted[i].klen, (int)sorted[i].vlen);
      tzo_writer_close(&w);
      raw = w.raw_len;
      part = w.comp_len;
      res->spilled_records += w.records;
    }
    res->output_bytes_with_overhead += raw; /* adjustSpillCounters, numSpills==0 branch :468-482 */
    res->index[p * 3 + 0] = seg_start;
    res->index[p * 3 + 1] = raw;
    res->index[p * 3 + 2] = part;
  }
  res->output_bytes_physical = (int64_t)res->file_out.len; /* fileOutputByteCounter :752 */
  tzo_spill_record_bytes(res->index, P, &res->index_out);
  if (sorted != m) free(sorted);
  free(m);
  return 0;
}

int tzo_pipelined_sort_fixed(const tzo_sorter_conf *conf, const uint8_t *kv, uint32_t klen, uint32_t vlen, uint64_t n,
                             tzo_sorter_result *res) {
  uint64_t *ko = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n ? n : 1));
  uint32_t *kl = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n ? n : 1));
  uint32_t *vl = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n ? n : 1));
  for (uint64_t i = 0; i < n; i++) { ko[i] = i * (uint64_t)(klen + vlen); kl[i] = klen; vl[i] = vlen; }
  int rc = tzo_pipelined_sort(conf, kv, ko, kl, vl, NULL, n, res);
  free(ko); free(kl); free(vl);
  return rc;
}

/* ------------------------------------------------------------------ UnorderedPartitionedKVWriter (no sort)
 * RL/common/writers/UnorderedPartitionedKVWriter.java: records are appended to a buffer and chained per partition
 * NEWEST FIRST (INDEX_NEXT = partitionPositions[p]; partitionPositions[p] = metaStart, :459-472); the final file walks
 * every partition's chain from its newest record back (writePartition :688-703, mergeAll :1058-1144) through an
 * IFile.Writer without run-length encoding (:1092).  Partitions without records are skipped: no bytes and an all-zero
 * index entry (TezSpillRecord starts zero-filled, SORT/TezSpillRecord.java:48-52).  This restates the case in which
 * everything fits ONE buffer (no spill thread ran): with several buffers / spills the record order inside a partition
 * depends on buffer arithmetic and thread timing, and parity is defined on the per-partition multiset. */
int tzo_unordered_write(const tzo_sorter_conf *conf, const uint8_t *kv, const uint64_t *key_off, const uint32_t *key_len,
                        const uint32_t *val_len, const int32_t *partition, uint64_t n, tzo_sorter_result *res) {
  memset(res, 0, sizeof(*res));
  tzo_buf_init(&res->file_out);
  tzo_buf_init(&res->index_out);
  const int P = conf->num_partitions;
  res->index = (int64_t *)calloc((size_t)P * 3, sizeof(int64_t));
  int32_t *part = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
  int64_t *next = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n ? n : 1));
  int64_t *head = (int64_t *)malloc(sizeof(int64_t) * (size_t)P);
  for (int p = 0; p < P; p++) head[p] = -1;
  for (uint64_t i = 0; i < n; i++) {
    int32_t p = partition ? partition[i]
                          : tzo_hash_partition(tzo_key_hash(conf->cmp_kind, kv + key_off[i], key_len[i]), P);
    if (p < 0 || p >= P) { free(part); free(next); free(head); return -1; }
    part[i] = p;
    next[i] = head[p];      /* INDEX_NEXT */
    head[p] = (int64_t)i;   /* partitionPositions[p] = this record */
    res->output_records++;
    res->output_bytes += (int64_t)key_len[i] + val_len[i];
  }
  for (int p = 0; p < P; p++) {
    if (head[p] < 0) continue;  /* "Skipping partition ... since it has no records" */
    tzo_ifile_writer w;
    size_t start = res->file_out.len;
    tzo_writer_open(&w, &res->file_out, 0);
    for (int64_t i = head[p]; i >= 0; i = next[i])
      tzo_writer_append(&w, kv + key_off[i], (int)key_len[i], kv + key_off[i] + key_len[i], (int)val_len[i]);
    tzo_writer_close(&w);
    res->index[3 * p] = (int64_t)start;
    res->index[3 * p + 1] = w.raw_len;
    res->index[3 * p + 2] = w.comp_len;
    res->output_bytes_with_overhead += w.raw_len;
  }
  res->output_bytes_physical = (int64_t)res->file_out.len;
  tzo_spill_record_bytes(res->index, P, &res->index_out);
  free(part); free(next); free(head);
  return 0;
}

void tzo_sorter_result_free(tzo_sorter_result *res) {
  tzo_buf_free(&res->file_out);
  tzo_buf_free(&res->index_out);
  free(res->index);
  res->index = NULL;
}

/* ------------------------------------------------------------------ TezMerger.MergeQueue */
typedef struct {
  tzo_ifile_reader rd;
  tzo_buf owned;      /* intermediate merged segments own their bytes */
  int is_owned;
  size_t length;      /* Segment.getLength() */
  int id;
} mseg;

typedef struct {
  mseg **heap;        /* 1-indexed, hadoop util.PriorityQueue */
  int size;
  int cmp_kind;
} mqueue;

static int mq_less(mqueue *q, mseg *a, mseg *b) {
  /* MergeQueue.lessThan :695-704 */
  return tzo_compare(q->cmp_kind, a->rd.key, a->rd.klen, b->rd.key, b->rd.klen) < 0;
}
static void mq_up(mqueue *q) {
  int i = q->size;
  mseg *node = q->heap[i];
  int j = i >> 1;
  while (j > 0 && mq_less(q, node, q->heap[j])) { q->heap[i] = q->heap[j]; i = j; j = j >> 1; }
  q->heap[i] = node;
}
static void mq_down(mqueue *q) {
  int i = 1;
  mseg *node = q->heap[i];
  int j = i << 1, k = j + 1;
  if (k <= q->size && mq_less(q, q->heap[k], q->heap[j])) j = k;
  while (j <= q->size && mq_less(q, q->heap[j], node)) {
    q->heap[i] = q->heap[j];
    i = j; j = i << 1; k = j + 1;
    if (k <= q->size && mq_less(q, q->heap[k], q->heap[j])) j = k;
  }
  q->heap[i] = node;
}
static void mq_put(mqueue *q, mseg *s) { q->size++; q->heap[q->size] = s; mq_up(q); }
static mseg *mq_top(mqueue *q) { return q->size > 0 ? q->heap[1] : NULL; }
static mseg *mq_pop(mqueue *q) {
  if (q->size <= 0) return NULL;
  mseg *r = q->heap[1];
  q->heap[1] = q->heap[q->size];
  q->heap[q->size] = NULL;
  q->size--;
  if (q->size > 0) mq_down(q);
  return r;
}

typedef struct {
  mqueue q;
  int check_same;
  int has_next;          /* KeyState or -1 for null */
  mseg *min_segment;
  tzo_buf prev_key;
  const uint8_t *key; int klen;
  const uint8_t *val; int vlen;
} miter;

static void mi_compare_with_next_top(miter *it, mseg *current) {
  /* compareKeyWithNextTopKey :641-652 */
  mseg *next_top = mq_top(&it->q);
  if (it->check_same && next_top != current && next_top != NULL) {
    int c = tzo_compare(it->q.cmp_kind, next_top->rd.key, next_top->rd.klen, it->prev_key.data, (int)it->prev_key.len);
    if (c == 0) it->has_next = TZO_SAME_KEY;
  }
}
static void mi_adjust(miter *it, mseg *reader) {
  /* adjustPriorityQueue :597-635 */
  if (it->check_same) {
    if (it->has_next == -1 || it->has_next != TZO_SAME_KEY) {
      it->prev_key.len = 0;
      tzo_buf_put(&it->prev_key, it->key, (size_t)it->klen); /* populatePreviousKey */
    }
  }
  it->has_next = tzo_reader_read_raw_key(&reader->rd);
  if (it->has_next == TZO_NEW_KEY) {
    mq_down(&it->q); /* adjustTop */
    mi_compare_with_next_top(it, reader);
  } else if (it->has_next == TZO_NO_KEY || it->has_next < 0) {
    it->has_next = TZO_NO_KEY;
    mq_pop(&it->q);
    mi_compare_with_next_top(it, NULL);
  }
}
static int mi_next(miter *it) {
  /* hasNext :1047-1063 + next :654-683 */
  if (it->q.size == 0) return 0;
  if (it->min_segment != NULL) {
    mi_adjust(it, it->min_segment);
    if (it->q.size == 0) { it->min_segment = NULL; return 0; }
  }
  it->min_segment = mq_top(&it->q);
  it->key = it->min_segment->rd.key;
  it->klen = it->min_segment->rd.klen;
  tzo_reader_next_raw_value(&it->min_segment->rd, &it->val, &it->vlen);
  return 1;
}
static int mi_is_same(miter *it) { return it->has_next == TZO_SAME_KEY; }

static int get_pass_factor(int factor, int pass_no, int num_segments) {
  /* getPassFactor :920-930 */
  if (pass_no > 1 || num_segments <= factor || factor == 1) return factor;
  int mod = (num_segments - 1) % (factor - 1);
  if (mod == 0) return factor;
  return mod + 1;
}
static int seg_len_cmp(const void *a, const void *b) {
  const mseg *x = *(mseg *const *)a, *y = *(mseg *const *)b;
  if (x->length == y->length) return x->id - y->id; /* List.sort is stable */
  return x->length < y->length ? -1 : 1;
}

/* when set, tzo_merge() only produces the writeFile output and the record count (no per-record stream): large merges */
static __thread int g_merge_ifile_only = 0;

int tzo_merge_ifile(const tzo_segment *segs, int nseg, int cmp_kind, int factor, int sort_segments, int check_for_same_keys,
                    int writer_rle, tzo_merge_result *res, double *seconds) {
  struct timespec t0, t1;
  g_merge_ifile_only = 1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int rc = tzo_merge(segs, nseg, cmp_kind, factor, sort_segments, check_for_same_keys, writer_rle, res);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  g_merge_ifile_only = 0;
  if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  return rc;
}

int tzo_merge(const tzo_segment *segs, int nseg, int cmp_kind, int factor, int sort_segments, int check_for_same_keys,
              int writer_rle, tzo_merge_result *res) {
  memset(res, 0, sizeof(*res));
  tzo_buf_init(&res->keys); tzo_buf_init(&res->vals); tzo_buf_init(&res->ifile);
  size_t cap = 1024;
  res->key_len = (uint32_t *)malloc(cap * 4);
  res->val_len = (uint32_t *)malloc(cap * 4);
  res->same_key = (uint8_t *)malloc(cap);

  int list_cap = nseg + 8;
  mseg **list = (mseg **)calloc((size_t)list_cap * 2, sizeof(mseg *));
  int nlist = 0, next_id = 0;
  for (int i = 0; i < nseg; i++) {
    mseg *s = (mseg *)calloc(1, sizeof(mseg));
    s->id = next_id++;
    s->length = segs[i].len;
    if (tzo_reader_open(&s->rd, segs[i].data, segs[i].len, segs[i].has_header, 1) != 0) { free(s); free(list); return -1; }
    list[nlist++] = s;
  }
  if (sort_segments) qsort(list, (size_t)nlist, sizeof(mseg *), seg_len_cmp); /* :541-543,570-572 */

  miter it;
  memset(&it, 0, sizeof(it));
  it.q.cmp_kind = cmp_kind;
  it.check_same = check_for_same_keys;
  tzo_buf_init(&it.prev_key);
  it.q.heap = (mseg **)calloc((size_t)(nseg + 2), sizeof(mseg *));

  tzo_ifile_writer fw;
  tzo_writer_open(&fw, &res->ifile, writer_rle);

  if (nlist > 0) {
    /* merge() :717-912 */
    int num_segments = nlist, orig_factor = factor, pass_no = 1, head = 0;
    for (;;) {
      factor = get_pass_factor(factor, pass_no, num_segments);
      mseg **to_merge = (mseg **)calloc((size_t)(factor + 1), sizeof(mseg *));
      int considered = 0, to_consider = factor;
      for (;;) {
        int take = to_consider;
        if (take > nlist - head) take = nlist - head;
        for (int t = 0; t < take; t++) {
          mseg *s = list[head++];
          int ks = tzo_reader_read_raw_key(&s->rd); /* segment.nextRawKey :770-788 */
          if (ks == TZO_NEW_KEY) { to_merge[considered++] = s; }
          else { num_segments--; if (s->is_owned) tzo_buf_free(&s->owned); free(s); }
        }
        if (considered == factor || head == nlist) break;
        to_consider = factor - considered;
      }
      it.q.size = 0;
      it.min_segment = NULL;
      it.has_next = -1;
      for (int t = 0; t < considered; t++) mq_put(&it.q, to_merge[t]);
      free(to_merge);

      if (num_segments <= factor) break; /* lazy final pass: iterate below */

      /* intermediate pass: writeFile(this, Writer(rle=false)) :859-875 ; re-insert by size :886-893 */
      mseg *tmp = (mseg *)calloc(1, sizeof(mseg));
      tmp->id = next_id++;
      tmp->is_owned = 1;
      tzo_buf_init(&tmp->owned);
      tzo_ifile_writer w;
      tzo_writer_open(&w, &tmp->owned, 0);
      mseg **consumed = (mseg **)calloc((size_t)considered + 1, sizeof(mseg *));
      int nconsumed = 0;
      for (int t = 1; t <= it.q.size; t++) consumed[nconsumed++] = it.q.heap[t];
      while (mi_next(&it)) {
        if (mi_is_same(&it)) tzo_writer_append(&w, NULL, 0, it.val, it.vlen);
        else tzo_writer_append(&w, it.key, it.klen, it.val, it.vlen);
      }
      tzo_writer_close(&w);
      for (int t = 0; t < nconsumed; t++) { if (consumed[t]->is_owned) tzo_buf_free(&consumed[t]->owned); free(consumed[t]); }
      free(consumed);
      tmp->length = tmp->owned.len;
      tzo_reader_open(&tmp->rd, tmp->owned.data, tmp->owned.len, 1, 1);
      res->passes++;
      /* compact remaining list and binary-search insert (Collections.binarySearch with segmentComparator) */
      int remain = nlist - head;
      memmove(list, list + head, sizeof(mseg *) * (size_t)remain);
      head = 0; nlist = remain;
      int lo = 0, hi = nlist - 1, posi = -1;
      while (lo <= hi) {
        int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1);
        size_t ml = list[mid]->length;
        if (ml < tmp->length) lo = mid + 1;
        else if (ml > tmp->length) hi = mid - 1;
        else { posi = mid; break; }
      }
      if (posi < 0) posi = lo;
      memmove(list + posi + 1, list + posi, sizeof(mseg *) * (size_t)(nlist - posi));
      list[posi] = tmp;
      nlist++;
      num_segments = nlist;
      pass_no++;
      factor = orig_factor;
    }

    /* final (lazy) pass: drain the iterator, record stream + writeFile */
    mseg **live = (mseg **)calloc((size_t)it.q.size + 1, sizeof(mseg *));
    int nlive = 0;
    for (int t = 1; t <= it.q.size; t++) live[nlive++] = it.q.heap[t];
    while (mi_next(&it)) {
      if (!g_merge_ifile_only && res->n == cap) {
        cap *= 2;
        res->key_len = (uint32_t *)realloc(res->key_len, cap * 4);
        res->val_len = (uint32_t *)realloc(res->val_len, cap * 4);
        res->same_key = (uint8_t *)realloc(res->same_key, cap);
      }
      int same = mi_is_same(&it);
      if (!g_merge_ifile_only) {
        res->key_len[res->n] = (uint32_t)it.klen;
        res->val_len[res->n] = (uint32_t)it.vlen;
        res->same_key[res->n] = (uint8_t)same;
        tzo_buf_put(&res->keys, it.key, (size_t)it.klen);
        tzo_buf_put(&res->vals, it.val, (size_t)it.vlen);
      }
      res->n++;
      /* TezMerger.writeFile :215-245 */
      if (same) tzo_writer_append(&fw, NULL, 0, it.val, it.vlen);
      else tzo_writer_append(&fw, it.key, it.klen, it.val, it.vlen);
    }
    for (int t = 0; t < nlive; t++) { if (live[t]->is_owned) tzo_buf_free(&live[t]->owned); free(live[t]); }
    free(live);
    for (int t = head; t < nlist; t++) { if (list[t]->is_owned) tzo_buf_free(&list[t]->owned); free(list[t]); }
  }
  tzo_writer_close(&fw);
  res->raw_len = fw.raw_len;
  res->comp_len = fw.comp_len;
  free(it.q.heap);
  tzo_buf_free(&it.prev_key);
  free(list);
  return 0;
}

void tzo_merge_result_free(tzo_merge_result *res) {
  tzo_buf_free(&res->keys); tzo_buf_free(&res->vals); tzo_buf_free(&res->ifile);
  free(res->key_len); free(res->val_len); free(res->same_key);
  res->key_len = res->val_len = NULL; res->same_key = NULL;
}

/* ------------------------------------------------------------------ generators */
uint64_t tzo_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
void tzo_gen_c2(uint8_t *dst, uint64_t first_index, uint64_t n, uint64_t seed) {
  /* key = 16B (two big-endian splitmix64 words), value = 64B (8 words), all functions of (seed, index) */
  for (uint64_t r = 0; r < n; r++) {
    uint64_t i = first_index + r;
    uint8_t *d = dst + r * 80;
    for (int w = 0; w < 10; w++) put_be64(d + 8 * w, tzo_splitmix64((seed << 56) ^ (i * 16 + (uint64_t)w)));
  }
}

/* sorted IFile segments of (Text word, 8-byte value = f(word)); words from a 2^24-id space, length
 * U[4,24] lower-case letters, both functions of the id alone; uncompressed, no run-length encoding in the inputs */
int tzo_c3_word(uint32_t id, uint8_t *out /* >= 24 bytes */) {
  uint64_t h = tzo_splitmix64(0xC3C3C3C3ull ^ ((uint64_t)id << 20));
  int len = 4 + (int)(h % 21);
  uint64_t r = 0;
  for (int i = 0; i < len; i++) {
    if ((i & 7) == 0) r = tzo_splitmix64(((uint64_t)id << 8) | (uint64_t)(i >> 3) | 0x5000000000ull);
    out[i] = (uint8_t)('a' + (r & 0xFF) % 26);
    r >>= 8;
  }
  return len;
}

typedef struct { uint8_t len; uint8_t w[24]; uint32_t id; } c3_rec;
static int c3_cmp(const void *a, const void *b) {
  const c3_rec *x = (const c3_rec *)a, *y = (const c3_rec *)b;
  int n = x->len < y->len ? x->len : y->len;
  int c = memcmp(x->w, y->w, (size_t)n);
  if (c) return c;
  return (int)x->len - (int)y->len;
}

void tzo_gen_c3_segment(uint64_t seed, uint32_t seg_index, uint64_t target_bytes, int id_bits, tzo_buf *out, uint64_t *nrecords) {
  const uint32_t id_mask = id_bits >= 32 ? 0xFFFFFFFFu : ((1u << id_bits) - 1u);
  uint64_t cap = target_bytes / 15 + 16, n = 0, bytes = 10;
  c3_rec *recs = (c3_rec *)malloc(cap * sizeof(c3_rec));
  for (uint64_t i = 0; bytes < target_bytes && n < cap; i++) {
    uint32_t id = (uint32_t)(tzo_splitmix64((seed << 56) ^ ((uint64_t)seg_index << 36) ^ i)) & id_mask;
    recs[n].id = id;
    recs[n].len = (uint8_t)tzo_c3_word(id, recs[n].w);
    bytes += 2u + 1u + recs[n].len + 8u;
    n++;
  }
  qsort(recs, n, sizeof(c3_rec), c3_cmp);
  /* every word at most once per segment (what a map-side combiner leaves): with duplicates INSIDE an unencoded segment
   * whose key also occurs in other segments, the reference's REPEAT_KEY placement depends on its heap's tie order
   * (parity unpinned) and the merged bytes would not be defined by the input alone */
  uint64_t m = 0;
  for (uint64_t i = 0; i < n; i++)
    if (m == 0 || c3_cmp(&recs[m - 1], &recs[i]) != 0) recs[m++] = recs[i];
  n = m;
  tzo_ifile_writer w;
  tzo_writer_open(&w, out, 0);
  uint8_t key[32], val[8];
  for (uint64_t i = 0; i < n; i++) {
    key[0] = recs[i].len;                       /* Text: vint(byte length) + UTF-8 */
    memcpy(key + 1, recs[i].w, recs[i].len);
    /* value = f(WORD), not f(id): several ids render the same short word, and equal keys must carry equal values for
     * the merged bytes to be independent of the order TezMerger's heap emits equal keys in (SURVEY 8c caveat ii) */
    uint64_t hv = 0xABCDull;
    for (int b = 0; b < recs[i].len; b++) hv = tzo_splitmix64(hv ^ recs[i].w[b]);
    put_be64(val, hv);
    tzo_writer_append(&w, key, 1 + recs[i].len, val, 8);
  }
  tzo_writer_close(&w);
  free(recs);
  if (nrecords) *nrecords = n;
}

typedef struct { uint64_t seed, target; uint32_t first, count, stride; int id_bits; tzo_buf *outs; uint64_t *nrec; } c3_job;
static void *c3_worker(void *arg) {
  c3_job *j = (c3_job *)arg;
  for (uint32_t s = j->first; s < j->count; s += j->stride) {
    tzo_buf_init(&j->outs[s]);
    tzo_gen_c3_segment(j->seed, s, j->target, j->id_bits, &j->outs[s], &j->nrec[s]);
  }
  return NULL;
}
void tzo_gen_c3_segments(uint64_t seed, uint32_t nseg, uint64_t target_bytes, int id_bits, int threads, tzo_buf *outs, uint64_t *nrec) {
  if (threads < 1) threads = 1;
  if ((uint32_t)threads > nseg) threads = (int)nseg;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
  c3_job *jobs = (c3_job *)calloc((size_t)threads, sizeof(c3_job));
  for (int t = 0; t < threads; t++) {
    jobs[t].seed = seed; jobs[t].target = target_bytes; jobs[t].first = (uint32_t)t; jobs[t].count = nseg;
    jobs[t].stride = (uint32_t)threads; jobs[t].id_bits = id_bits; jobs[t].outs = outs; jobs[t].nrec = nrec;
    pthread_create(&th[t], NULL, c3_worker, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

/* ------------------------------------------------------------------ CPU baseline driver */
typedef struct {
  const tzo_sorter_conf *conf;
  const uint8_t *kv;
  uint32_t klen, vlen;
  uint64_t n;
  uint64_t out_bytes;
  int sort_threads;
} bench_task;

static void *bench_worker(void *arg) {
  bench_task *t = (bench_task *)arg;
  tzo_sorter_result r;
  tzo_sorter_conf c = *t->conf;
  c.sort_threads = t->sort_threads;
  tzo_pipelined_sort_fixed(&c, t->kv, t->klen, t->vlen, t->n, &r);
  t->out_bytes = r.file_out.len;
  tzo_sorter_result_free(&r);
  return NULL;
}

double tzo_bench_pipelined_fixed(const tzo_sorter_conf *conf, const uint8_t *kv, uint32_t klen, uint32_t vlen, uint64_t n,
                                 int tasks, uint64_t *out_bytes) {
  if (tasks < 1) tasks = 1;
  bench_task *bt = (bench_task *)calloc((size_t)tasks, sizeof(bench_task));
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)tasks);
  uint64_t per = n / (uint64_t)tasks;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < tasks; t++) {
    bt[t].conf = conf;
    bt[t].kv = kv + (uint64_t)t * per * (klen + vlen);
    bt[t].klen = klen; bt[t].vlen = vlen;
    /* one task: the sorter's own span-sort pool (conf->sort_threads); many concurrent tasks: one sort thread each */
    bt[t].sort_threads = tasks == 1 ? (conf->sort_threads > 0 ? conf->sort_threads : 1) : 1;
    bt[t].n = (t == tasks - 1) ? n - per * (uint64_t)(tasks - 1) : per;
    pthread_create(&th[t], NULL, bench_worker, &bt[t]);
  }
  uint64_t ob = 0;
  for (int t = 0; t < tasks; t++) { pthread_join(th[t], NULL); ob += bt[t].out_bytes; }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (out_bytes) *out_bytes = ob;
  free(bt); free(th);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
