// common.cuh -- shared device/host helpers for libtezgpu (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <string>

namespace tezgpu {

struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define TG_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      throw ::tezgpu::Error(_e == cudaErrorMemoryAllocation ? -3 : -2,                                  \
                            std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                                std::to_string(__LINE__) + ")");                                        \
    }                                                                                                   \
  } while (0)

#define TG_CHECK(cond, code, msg)                      \
  do {                                                 \
    if (!(cond)) throw ::tezgpu::Error((code), (msg)); \
  } while (0)

static inline uint64_t div_up(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
static inline uint64_t align_up(uint64_t a, uint64_t b) { return div_up(a, b) * b; }

// comparator ids (include/tezgpu.h)
enum { CMP_BYTES = 0, CMP_TEXT = 1, CMP_BYTESWRITABLE = 2, CMP_INT = 3, CMP_LONG = 4 };

// -------------------------------------------------------------------------------------------- device helpers
// hadoop WritableUtils.decodeVIntSize on the first byte
__host__ __device__ __forceinline__ int vint_decode_size(uint8_t first) {
  int v = (int)(int8_t)first;
  if (v >= -112) return 1;
  if (v < -120) return -119 - v;
  return -111 - v;
}
// WritableUtils.getVIntSize for non-negative lengths
__host__ __device__ __forceinline__ int vint_size_u32(uint32_t v) {
  if (v <= 127) return 1;
  if (v < (1u << 8)) return 2;
  if (v < (1u << 16)) return 3;
  if (v < (1u << 24)) return 4;
  return 5;
}
// byte `b` (0-based) of the vint encoding of non-negative v
__host__ __device__ __forceinline__ uint8_t vint_byte_u32(uint32_t v, int b) {
  int sz = vint_size_u32(v);
  if (sz == 1) return (uint8_t)v;
  if (b == 0) return (uint8_t)(-112 - (sz - 1));  // -113..-116
  int shift = (sz - 1 - b) * 8;
  return (uint8_t)(v >> shift);
}

// bytes of key content to skip before comparing / hashing (Text: vint prefix, BytesWritable: 4-byte length)
__device__ __forceinline__ uint32_t key_content_skip(int cmp, const uint8_t *key, uint32_t klen) {
  if (klen == 0) return 0;
  if (cmp == CMP_TEXT) {
    uint32_t s = (uint32_t)vint_decode_size(key[0]);
    return s < klen ? s : klen;
  }
  if (cmp == CMP_BYTESWRITABLE) return klen < 4 ? klen : 4;
  return 0;
}

// normalised content byte i: unsigned lexicographic order over these bytes == comparator order
// (IntWritable / LongWritable: two's complement big-endian => flip the sign bit of byte 0)
__device__ __forceinline__ uint32_t norm_byte(int cmp, const uint8_t *content, uint32_t i) {
  uint32_t b = content[i];
  if (i == 0 && (cmp == CMP_INT || cmp == CMP_LONG)) b ^= 0x80u;
  return b;
}

// WritableComparator.hashBytes
__device__ __forceinline__ int32_t hash_bytes_dev(const uint8_t *p, uint32_t n) {
  uint32_t h = 1;
  for (uint32_t i = 0; i < n; i++) h = 31u * h + (uint32_t)(int32_t)(int8_t)p[i];
  return (int32_t)h;
}

// key.hashCode() for the supported key classes
__device__ __forceinline__ int32_t key_hash_dev(int cmp, const uint8_t *key, uint32_t klen) {
  if (cmp == CMP_INT && klen >= 4)
    return (int32_t)(((uint32_t)key[0] << 24) | ((uint32_t)key[1] << 16) | ((uint32_t)key[2] << 8) | key[3]);
  if (cmp == CMP_LONG && klen >= 8) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | key[i];
    return (int32_t)(uint32_t)(v ^ (v >> 32));
  }
  uint32_t s = key_content_skip(cmp, key, klen);
  return hash_bytes_dev(key + s, klen - s);
}

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t *p, uint64_t v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace tezgpu
