// device_util.h -- host-side RAII helpers (device / pinned buffers) for libtezgpu.
#pragma once
#include "common.cuh"

namespace tezgpu {

struct DeviceBuffer {
  void *p = nullptr;
  size_t cap = 0;
  ~DeviceBuffer() { release(); }
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer &) = delete;
  DeviceBuffer &operator=(const DeviceBuffer &) = delete;
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  // grow-only; contents are NOT preserved
  void ensure(size_t bytes) {
    bytes = align_up(bytes ? bytes : 16, 256);
    if (bytes <= cap) return;
    release();
    TG_CUDA(cudaMalloc(&p, bytes));
    cap = bytes;
  }
  // grow preserving the first `keep` bytes
  void grow_preserve(size_t bytes, size_t keep, cudaStream_t st) {
    bytes = align_up(bytes ? bytes : 16, 256);
    if (bytes <= cap) return;
    size_t ncap = cap ? cap : 4096;
    while (ncap < bytes) ncap = ncap + ncap / 2;
    ncap = align_up(ncap, 256);
    void *np = nullptr;
    cudaError_t e = cudaMalloc(&np, ncap);
    if (e != cudaSuccess) {
      cudaGetLastError();
      ncap = bytes;  // retry with the exact size before giving up
      TG_CUDA(cudaMalloc(&np, ncap));
    }
    if (keep) {
      TG_CUDA(cudaMemcpyAsync(np, p, keep, cudaMemcpyDeviceToDevice, st));
      TG_CUDA(cudaStreamSynchronize(st));
    }
    if (p) cudaFree(p);
    p = np;
    cap = ncap;
  }
  template <typename T>
  T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinnedBuffer {
  void *p = nullptr;
  size_t cap = 0;
  ~PinnedBuffer() { release(); }
  PinnedBuffer() = default;
  PinnedBuffer(const PinnedBuffer &) = delete;
  PinnedBuffer &operator=(const PinnedBuffer &) = delete;
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  void ensure(size_t bytes) {
    bytes = align_up(bytes ? bytes : 16, 4096);
    if (bytes <= cap) return;
    release();
    TG_CUDA(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
    cap = bytes;
  }
  template <typename T>
  T *as() const { return reinterpret_cast<T *>(p); }
};

struct EventTimer {
  cudaEvent_t ev[8];
  int n = 0;
  EventTimer() { for (auto &e : ev) cudaEventCreate(&e); }
  ~EventTimer() { for (auto &e : ev) cudaEventDestroy(e); }
  void mark(cudaStream_t st) { if (n < 8) cudaEventRecord(ev[n++], st); }
  float ms(int a, int b) {
    float f = 0;
    if (a < n && b < n) cudaEventElapsedTime(&f, ev[a], ev[b]);
    return f;
  }
  void reset() { n = 0; }
};

}  // namespace tezgpu
