// merger_api.inl -- extern "C" merge entry points (include/tezgpu.h)
struct tezgpu_merger {
  Merger m;
  explicit tezgpu_merger(const tezgpu_conf &c) : m(c) {}
};

extern "C" {

int32_t tezgpu_merge_open(const tezgpu_conf *conf, const tezgpu_segment *segs, uint32_t nseg, tezgpu_merger **out) {
  tezgpu_merger *h = nullptr;
  try {
    TG_CHECK(conf && out && (segs || nseg == 0), TEZGPU_E_INVALID, "null argument");
    TG_CHECK(conf->abi_version == TEZGPU_ABI_VERSION, TEZGPU_E_INVALID, "tezgpu_conf.abi_version mismatch");
    h = new tezgpu_merger(*conf);
    h->m.open(segs, nseg);
    *out = h;
  } catch (const tezgpu::Error &e) {
    delete h;
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception &e) {
    delete h;
    g_last_error = e.what();
    return TEZGPU_E_INVALID;
  }
  return TEZGPU_OK;
}

int32_t tezgpu_merge_reopen(tezgpu_merger *m, const tezgpu_segment *segs, uint32_t nseg) {
  TG_API_BEGIN
  TG_CHECK(m && (segs || nseg == 0), TEZGPU_E_INVALID, "null argument");
  m->m.launches = 0;
  m->m.open(segs, nseg);
  TG_API_END
}

int32_t tezgpu_merge_set_check_for_same_keys(tezgpu_merger *m, int32_t check_for_same_keys) {
  TG_API_BEGIN
  TG_CHECK(m, TEZGPU_E_INVALID, "null handle");
  m->m.pipe.merge_check_same = check_for_same_keys ? 1 : 0;
  TG_API_END
}

int32_t tezgpu_merge_parse_info(tezgpu_merger *m, int32_t *mode, int32_t *rounds) {
  TG_API_BEGIN
  TG_CHECK(m, TEZGPU_E_INVALID, "null handle");
  if (mode) *mode = m->m.parse_mode;
  if (rounds) *rounds = m->m.parse_rounds;
  TG_API_END
}

int32_t tezgpu_merge_counts(tezgpu_merger *m, uint64_t *records, uint64_t *kv_bytes) {
  TG_API_BEGIN
  TG_CHECK(m, TEZGPU_E_INVALID, "null handle");
  if (records) *records = m->m.n;
  if (kv_bytes) *kv_bytes = m->m.kv_bytes;
  TG_API_END
}

int32_t tezgpu_merge_next_batch(tezgpu_merger *m, uint8_t *out_kv, uint64_t cap, tezgpu_kv_index *idx, uint32_t idx_cap,
                                uint32_t *n) {
  TG_API_BEGIN
  TG_CHECK(m && out_kv && idx && n, TEZGPU_E_INVALID, "null argument");
  m->m.next_batch(out_kv, cap, idx, idx_cap, n);
  TG_API_END
}

uint64_t tezgpu_merge_output_bound(const tezgpu_merger *m) { return m ? m->m.output_bound() : 0; }

int32_t tezgpu_merge_write_ifile_device(tezgpu_merger *m, void *d_out, uint64_t out_cap, int32_t rle, int64_t *raw_len,
                                        int64_t *part_len, tezgpu_stats *stats) {
  TG_API_BEGIN
  TG_CHECK(m && d_out, TEZGPU_E_INVALID, "null argument");
  m->m.write_device((uint8_t *)d_out, out_cap, rle, raw_len, part_len, stats);
  TG_API_END
}

int32_t tezgpu_merge_write_ifile(tezgpu_merger *m, const char *path, uint8_t *out, uint64_t out_cap, int32_t rle,
                                 int64_t *raw_len, int64_t *part_len, tezgpu_stats *stats) {
  TG_API_BEGIN
  TG_CHECK(m && (path || out), TEZGPU_E_INVALID, "null argument");
  Merger &mm = m->m;
  uint64_t bound = mm.output_bound();
  mm.d_out.ensure(bound);
  int64_t raw = 0, part = 0;
  mm.write_device(mm.d_out.as<uint8_t>(), mm.d_out.cap, rle, &raw, &part, stats);
  uint8_t *host = out;
  if (!host) {
    mm.h_out.ensure((size_t)part + 16);
    host = mm.h_out.as<uint8_t>();
  } else {
    TG_CHECK((uint64_t)part <= out_cap, TEZGPU_E_NOMEM, "output buffer too small for the merged segment");
  }
  TG_CUDA(cudaMemcpyAsync(host, mm.d_out.p, (size_t)part, cudaMemcpyDeviceToHost, mm.pipe.stream));
  TG_CUDA(cudaStreamSynchronize(mm.pipe.stream));
  if (path) write_file_0640(path, host, (size_t)part);
  if (raw_len) *raw_len = raw;
  if (part_len) *part_len = part;
  TG_API_END
}

int32_t tezgpu_merge_write_partitions_device(tezgpu_merger *m, void *d_out, uint64_t out_cap, int32_t rle,
                                             uint64_t *out_len, int64_t *index, tezgpu_stats *stats) {
  TG_API_BEGIN
  TG_CHECK(m && d_out, TEZGPU_E_INVALID, "null argument");
  m->m.write_partitions_device((uint8_t *)d_out, out_cap, rle, out_len, index, stats);
  TG_API_END
}

int32_t tezgpu_merge_write_partitions(tezgpu_merger *m, const char *out_path, const char *index_path, int32_t rle,
                                      int64_t *index, tezgpu_stats *stats) {
  TG_API_BEGIN
  TG_CHECK(m && out_path && index_path, TEZGPU_E_INVALID, "null argument");
  Merger &mm = m->m;
  const int P = mm.pipe.conf.num_partitions;
  mm.d_out.ensure(mm.output_bound());
  std::vector<int64_t> idx((size_t)P * 3, 0);
  uint64_t len = 0;
  tezgpu_stats st;
  mm.write_partitions_device(mm.d_out.as<uint8_t>(), mm.d_out.cap, rle, &len, idx.data(), &st);
  mm.h_out.ensure(len + 16);
  if (len) {
    TG_CUDA(cudaMemcpyAsync(mm.h_out.p, mm.d_out.p, len, cudaMemcpyDeviceToHost, mm.pipe.stream));
    TG_CUDA(cudaStreamSynchronize(mm.pipe.stream));
  }
  write_file_0640(out_path, mm.h_out.p, len);
  std::vector<uint8_t> b;
  spill_record_bytes(idx.data(), P, b);
  write_file_0640(index_path, b.data(), b.size());
  if (index) memcpy(index, idx.data(), (size_t)P * 24);
  if (stats) *stats = st;
  TG_API_END
}

void *tezgpu_merge_stream(tezgpu_merger *m) { return m ? (void *)m->m.pipe.stream : nullptr; }

int32_t tezgpu_merge_close(tezgpu_merger *m) {
  TG_API_BEGIN
  delete m;
  TG_API_END
}

}  // extern "C"
