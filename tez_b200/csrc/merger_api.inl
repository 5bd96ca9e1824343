// merger_api.inl -- extern "C" merge entry points
extern "C" {
int32_t tezgpu_merge_open(const tezgpu_conf *, const tezgpu_segment *, uint32_t, tezgpu_merger **) { g_last_error = "merge not built"; return TEZGPU_E_UNSUPPORTED; }
int32_t tezgpu_merge_counts(tezgpu_merger *, uint64_t *, uint64_t *) { return TEZGPU_E_UNSUPPORTED; }
int32_t tezgpu_merge_next_batch(tezgpu_merger *, uint8_t *, uint64_t, tezgpu_kv_index *, uint32_t, uint32_t *) { return TEZGPU_E_UNSUPPORTED; }
int32_t tezgpu_merge_write_ifile(tezgpu_merger *, const char *, uint8_t *, uint64_t, int32_t, int64_t *, int64_t *, tezgpu_stats *) { return TEZGPU_E_UNSUPPORTED; }
uint64_t tezgpu_merge_output_bound(const tezgpu_merger *) { return 0; }
int32_t tezgpu_merge_write_ifile_device(tezgpu_merger *, void *, uint64_t, int32_t, int64_t *, int64_t *, tezgpu_stats *) { return TEZGPU_E_UNSUPPORTED; }
void *tezgpu_merge_stream(tezgpu_merger *) { return nullptr; }
int32_t tezgpu_merge_close(tezgpu_merger *) { return TEZGPU_OK; }
}
