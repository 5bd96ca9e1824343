// shuffle_wire.inl -- SURVEY 8 f-2: the wire format between ShuffleHandler and FetcherOrderedGrouped, so that map
// outputs that live in HBM can be served to consumers that are NOT on the NVLink domain (and unmodified fetchers).
//
// A response to /mapOutput?job=..&dag=..&reduce=R&map=M1,M2,.. (RL/common/shuffle/ShuffleUtils.java:209-249) is, per
// requested map output: a ShuffleHeader (OG/ShuffleHeader.java:101-106: Text.writeString(mapId), vlong compressedLength,
// vlong uncompressedLength, vint forReduce) followed by compressedLength bytes -- the partition's IFile segment exactly as
// it sits in file.out (FetcherOrderedGrouped.copyMapOutput reads the header, reserves, then reads that many bytes:
// OG/FetcherOrderedGrouped.java:437-632).  compressedLength = TezIndexRecord.partLength, uncompressedLength = rawLength.
//
// Host code only (included by tezgpu_api.cu inside extern "C"): the segment bytes come out of the device-resident
// file.out with one cudaMemcpyAsync per partition, straight behind their header in the caller's (pinned) buffer.

static inline int wire_vlong_size(int64_t v) {
  if (v >= -112 && v <= 127) return 1;
  uint64_t u = v < 0 ? ~(uint64_t)v : (uint64_t)v;
  int n = 0;
  while (u) { u >>= 8; n++; }
  return n + 1;
}
// WritableUtils.writeVLong
static inline int wire_put_vlong(uint8_t *out, int64_t v) {
  if (v >= -112 && v <= 127) { out[0] = (uint8_t)(int8_t)v; return 1; }
  int len = -112;
  uint64_t u = (uint64_t)v;
  if (v < 0) { u = ~u; len = -120; }
  int n = 0;
  for (uint64_t t = u; t; t >>= 8) n++;
  out[0] = (uint8_t)(int8_t)(len - n);
  for (int i = 0; i < n; i++) out[1 + i] = (uint8_t)(u >> (8 * (n - 1 - i)));
  return n + 1;
}
// WritableUtils.readVLong; returns bytes consumed or 0 when the input is too short / malformed
static inline int wire_get_vlong(const uint8_t *in, uint64_t avail, int64_t *v) {
  if (avail < 1) return 0;
  const int8_t first = (int8_t)in[0];
  if (first >= -112) { *v = first; return 1; }
  const int n = vint_decode_size((uint8_t)first);
  if ((uint64_t)n > avail) return 0;
  uint64_t u = 0;
  for (int i = 1; i < n; i++) u = (u << 8) | in[i];
  *v = first < -120 ? (int64_t)~u : (int64_t)u;
  return n;
}

uint64_t tezgpu_shuffle_header_size(const char *map_id, int64_t part_len, int64_t raw_len, int32_t reduce) {
  const size_t idlen = map_id ? strlen(map_id) : 0;
  return (uint64_t)wire_vlong_size((int64_t)idlen) + idlen + (uint64_t)wire_vlong_size(part_len) + (uint64_t)wire_vlong_size(raw_len) +
         (uint64_t)wire_vlong_size(reduce);
}

int32_t tezgpu_shuffle_header_write(const char *map_id, int64_t part_len, int64_t raw_len, int32_t reduce, uint8_t *out, uint64_t cap,
                                    uint64_t *len) {
  TG_API_BEGIN
  TG_CHECK(map_id && out && len, TEZGPU_E_INVALID, "null argument");
  const uint64_t need = tezgpu_shuffle_header_size(map_id, part_len, raw_len, reduce);
  TG_CHECK(need <= cap, TEZGPU_E_NOMEM, "buffer too small for the shuffle header");
  const size_t idlen = strlen(map_id);
  uint64_t o = 0;
  o += (uint64_t)wire_put_vlong(out + o, (int64_t)idlen);   // Text.writeString: vint(byte length) + UTF-8 bytes
  memcpy(out + o, map_id, idlen);
  o += idlen;
  o += (uint64_t)wire_put_vlong(out + o, part_len);
  o += (uint64_t)wire_put_vlong(out + o, raw_len);
  o += (uint64_t)wire_put_vlong(out + o, reduce);
  *len = o;
  TG_API_END
}

int32_t tezgpu_shuffle_header_read(const uint8_t *in, uint64_t avail, char *map_id, uint64_t map_id_cap, int64_t *part_len,
                                   int64_t *raw_len, int32_t *reduce, uint64_t *consumed) {
  TG_API_BEGIN
  TG_CHECK(in && map_id && part_len && raw_len && reduce && consumed, TEZGPU_E_INVALID, "null argument");
  uint64_t o = 0;
  int64_t idlen = 0, r = 0;
  int n = wire_get_vlong(in, avail, &idlen);
  TG_CHECK(n > 0, TEZGPU_E_FORMAT, "truncated shuffle header");
  o += (uint64_t)n;
  // WritableUtils.readStringSafely(in, MAX_ID_LENGTH = 1000) (OG/ShuffleHeader.java:48,83)
  TG_CHECK(idlen >= 0 && idlen <= 1000, TEZGPU_E_FORMAT, "shuffle header: map id length out of range");
  TG_CHECK(o + (uint64_t)idlen <= avail, TEZGPU_E_FORMAT, "truncated shuffle header");
  TG_CHECK((uint64_t)idlen + 1 <= map_id_cap, TEZGPU_E_NOMEM, "map id buffer too small");
  memcpy(map_id, in + o, (size_t)idlen);
  map_id[idlen] = 0;
  o += (uint64_t)idlen;
  n = wire_get_vlong(in + o, avail - o, part_len);
  TG_CHECK(n > 0, TEZGPU_E_FORMAT, "truncated shuffle header");
  o += (uint64_t)n;
  n = wire_get_vlong(in + o, avail - o, raw_len);
  TG_CHECK(n > 0, TEZGPU_E_FORMAT, "truncated shuffle header");
  o += (uint64_t)n;
  n = wire_get_vlong(in + o, avail - o, &r);
  TG_CHECK(n > 0, TEZGPU_E_FORMAT, "truncated shuffle header");
  o += (uint64_t)n;
  TG_CHECK(*part_len >= 0 && *raw_len >= 0 && r >= 0 && r <= 0x7fffffffll, TEZGPU_E_FORMAT, "shuffle header: negative length or partition");
  *reduce = (int32_t)r;
  *consumed = o;
  TG_API_END
}

uint64_t tezgpu_shuffle_serve_bound(const char *map_id, const int64_t *index, int32_t reduce0, int32_t nreduce) {
  uint64_t total = 0;
  for (int32_t p = reduce0; p < reduce0 + nreduce; p++)
    total += tezgpu_shuffle_header_size(map_id, index[3 * p + 2], index[3 * p + 1], p) + (uint64_t)index[3 * p + 2];
  return total;
}

int32_t tezgpu_shuffle_serve(int32_t device, const void *d_file_out, const int64_t *index, const char *map_id, int32_t reduce0,
                             int32_t nreduce, uint8_t *out, uint64_t cap, uint64_t *len, void *stream) {
  TG_API_BEGIN
  TG_CHECK(d_file_out && index && map_id && out && len && reduce0 >= 0 && nreduce >= 0, TEZGPU_E_INVALID, "bad argument");
  TG_CHECK(tezgpu_shuffle_serve_bound(map_id, index, reduce0, nreduce) <= cap, TEZGPU_E_NOMEM, "buffer too small for the shuffle response");
  TG_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  uint64_t o = 0;
  for (int32_t p = reduce0; p < reduce0 + nreduce; p++) {
    const int64_t start = index[3 * p], raw = index[3 * p + 1], part = index[3 * p + 2];
    TG_CHECK(start >= 0 && raw >= 0 && part >= 0, TEZGPU_E_INVALID, "negative spill index entry");
    uint64_t hl = 0;
    int32_t rc = tezgpu_shuffle_header_write(map_id, part, raw, p, out + o, cap - o, &hl);
    TG_CHECK(rc == TEZGPU_OK, rc, g_last_error);
    o += hl;
    if (part) TG_CUDA(cudaMemcpyAsync(out + o, (const uint8_t *)d_file_out + start, (size_t)part, cudaMemcpyDeviceToHost, st));
    o += (uint64_t)part;
  }
  TG_CUDA(cudaStreamSynchronize(st));
  *len = o;
  TG_API_END
}

int32_t tezgpu_shuffle_receive(const uint8_t *in, uint64_t len, tezgpu_wire_segment *segs, uint32_t cap, uint32_t *n) {
  TG_API_BEGIN
  TG_CHECK(in && n && (segs || cap == 0), TEZGPU_E_INVALID, "null argument");
  uint64_t o = 0;
  uint32_t k = 0;
  while (o < len) {
    tezgpu_wire_segment s;
    memset(&s, 0, sizeof(s));
    uint64_t used = 0;
    int32_t rc = tezgpu_shuffle_header_read(in + o, len - o, s.map_id, sizeof(s.map_id), &s.part_len, &s.raw_len, &s.reduce, &used);
    TG_CHECK(rc == TEZGPU_OK, rc, g_last_error);
    o += used;
    TG_CHECK((uint64_t)s.part_len <= len - o, TEZGPU_E_FORMAT, "shuffle response ends inside a segment");
    s.offset = o;
    o += (uint64_t)s.part_len;
    if (k < cap) segs[k] = s;
    k++;
  }
  *n = k;
  TG_CHECK(k <= cap, TEZGPU_E_NOMEM, "more segments in the response than the table holds");
  TG_API_END
}
