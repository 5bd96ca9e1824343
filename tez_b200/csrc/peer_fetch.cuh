// peer_fetch.cuh -- the shuffle transfer on an NVLink/NVSwitch box: the consumer pulls byte ranges of the producers'
// file.out straight out of their HBM (CUDA IPC peer mappings) with every SM, instead of the HTTP round trip of
// ShuffleHandler + FetcherOrderedGrouped.copyMapOutput (OG/FetcherOrderedGrouped.java:437-632).
// One launch moves any number of (source, destination, length) ranges; ranges whose two addresses agree modulo 16
// move as 128-bit words (the caller picks destinations that way), anything else falls back to bytes.
#pragma once
#include "common.cuh"

namespace tezgpu {

struct FetchRange {
  const uint8_t *src;
  uint8_t *dst;
  uint64_t len;
  uint64_t chunk0;  // index of this range's first chunk in the launch (prefix sum, filled by the host)
};

constexpr uint32_t FETCH_THREADS = 512;
constexpr uint32_t FETCH_UNROLL = 8;
constexpr uint64_t FETCH_CHUNK = (uint64_t)FETCH_THREADS * FETCH_UNROLL * 16 * 2;  // 128 KiB of 16-byte words per chunk

// the producer rewrites its buffer every step: never serve a peer byte from a local cache line (ld.cv)
__device__ __forceinline__ uint4 ld_peer_16(const uint4 *p) { return __ldcv(p); }
__device__ __forceinline__ uint8_t ld_peer_8(const uint8_t *p) { return __ldcv(p); }  // head / tail / misaligned bytes: same rule

__device__ __forceinline__ void fetch_ranges_body(const FetchRange *__restrict__ ranges, uint32_t nranges, uint64_t nchunks) {
  for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    uint32_t lo = 0, hi = nranges;  // last range with chunk0 <= c
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (ranges[mid].chunk0 <= c) lo = mid; else hi = mid;
    }
    const FetchRange r = ranges[lo];
    const uint64_t k = c - r.chunk0;
    const uint32_t mis = (uint32_t)((uintptr_t)r.src & 15u);
    if (mis != (uint32_t)((uintptr_t)r.dst & 15u)) {  // incompatible alignment: bytes
      const uint64_t a = k * FETCH_CHUNK, b = min(r.len, a + FETCH_CHUNK);
      for (uint64_t i = a + threadIdx.x; i < b; i += FETCH_THREADS) r.dst[i] = ld_peer_8(r.src + i);
      continue;
    }
    // body: the 16-byte words between the first and the last aligned address of the range
    const uint64_t head = min(r.len, (uint64_t)((16u - mis) & 15u));
    const uint64_t words = (r.len - head) >> 4;
    const uint4 *s16 = reinterpret_cast<const uint4 *>(r.src + head);
    uint4 *d16 = reinterpret_cast<uint4 *>(r.dst + head);
    const uint64_t w0 = k * (FETCH_CHUNK / 16), w1 = min(words, w0 + FETCH_CHUNK / 16);
    uint64_t i = w0 + threadIdx.x;
    for (; i + (uint64_t)(FETCH_UNROLL - 1) * FETCH_THREADS < w1; i += (uint64_t)FETCH_UNROLL * FETCH_THREADS) {
      uint4 v[FETCH_UNROLL];
#pragma unroll
      for (uint32_t u = 0; u < FETCH_UNROLL; u++) v[u] = ld_peer_16(s16 + i + (uint64_t)u * FETCH_THREADS);
#pragma unroll
      for (uint32_t u = 0; u < FETCH_UNROLL; u++) d16[i + (uint64_t)u * FETCH_THREADS] = v[u];
    }
    for (; i < w1; i += FETCH_THREADS) d16[i] = ld_peer_16(s16 + i);
    if (k == 0) {  // the unaligned head and tail bytes of the range travel with its first chunk
      const uint64_t tail0 = head + (words << 4);
      if (threadIdx.x < head) r.dst[threadIdx.x] = ld_peer_8(r.src + threadIdx.x);
      if (threadIdx.x >= 32 && tail0 + (threadIdx.x - 32) < r.len && threadIdx.x - 32 < 16)
        r.dst[tail0 + (threadIdx.x - 32)] = ld_peer_8(r.src + tail0 + (threadIdx.x - 32));
    }
  }
}

__global__ void __launch_bounds__(FETCH_THREADS)
    k_fetch_ranges(const FetchRange *__restrict__ ranges, uint32_t nranges, uint64_t nchunks) {
  fetch_ranges_body(ranges, nranges, nchunks);
}

// the usual case -- one range per peer GPU -- needs no device-side table: the ranges ride in the parameter space
constexpr uint32_t FETCH_INLINE_RANGES = 16;
struct FetchRangeList { FetchRange r[FETCH_INLINE_RANGES]; };
__global__ void __launch_bounds__(FETCH_THREADS)
    k_fetch_ranges_inline(const __grid_constant__ FetchRangeList lst, uint32_t nranges, uint64_t nchunks) {
  fetch_ranges_body(lst.r, nranges, nchunks);
}

// number of chunks a range of `len` bytes starting at `src` occupies (at least one, so head/tail bytes always move)
static inline uint64_t fetch_chunks(const void *src, const void *dst, uint64_t len) {
  const uint32_t mis = (uint32_t)((uintptr_t)src & 15u);
  if (mis != (uint32_t)((uintptr_t)dst & 15u)) return len ? div_up(len, FETCH_CHUNK) : 0;
  if (!len) return 0;
  const uint64_t head = std::min<uint64_t>(len, (16u - mis) & 15u);
  const uint64_t words = (len - head) >> 4;
  return std::max<uint64_t>(1, div_up(words * 16, FETCH_CHUNK));
}

}  // namespace tezgpu
