// Shared host-side helpers for libcslam_b200: error reporting across the C ABI,
// launch accounting, and small device utilities used by several kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../../include/cslam_b200.h"

namespace cslam {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define CSLAM_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t _e = (call);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::cslam::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                \
                         cudaGetErrorString(_e));                                     \
      return (_e == cudaErrorMemoryAllocation) ? CSLAM_ERR_OOM : CSLAM_ERR_CUDA;      \
    }                                                                                 \
  } while (0)

#define CSLAM_TRY(call)            \
  do {                             \
    int _s = (call);               \
    if (_s != CSLAM_OK) return _s; \
  } while (0)

#define CSLAM_REQUIRE(cond, ...)         \
  do {                                   \
    if (!(cond)) {                       \
      ::cslam::set_error(__VA_ARGS__);   \
      return CSLAM_ERR_INVALID;          \
    }                                    \
  } while (0)

// Check for launch-configuration errors right after a <<<>>> launch.
#define CSLAM_LAUNCH_CHECK()        \
  do {                              \
    ::cslam::count_launch();        \
    CSLAM_CUDA(cudaGetLastError()); \
  } while (0)

// RAII device-context guard: every ABI call switches to the handle's device.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Device allocations.  Buffers up to kPoolMaxBytes come from the device's stream-ordered memory pool
// with its release threshold raised to "never": a handle that is created and destroyed per call
// (cslam_mac: ~40 buffers per selection) then recycles its memory instead of going to the driver,
// where a cudaFree / cudaMalloc that trims or maps memory was observed to stall for up to a second.
// Semantics stay those of cudaMalloc / cudaFree: the memory is usable on every stream when
// dev_alloc returns, and dev_free waits for the device before the buffer can be reused.
// CSLAM_DEV_POOL=0 switches the pool off.  Larger buffers (keyframe pools) use cudaMalloc / cudaFree.
constexpr size_t kPoolMaxBytes = size_t(64) << 20;
int pool_alloc(void** p, size_t bytes);   // lib.cu; CSLAM_OK, or a status with *p = nullptr (caller falls back)
bool pool_free(void* p);                  // true: p came from the pool and has been returned to it

template <typename T>
inline int dev_alloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  const size_t bytes = count * sizeof(T);
  if (bytes <= kPoolMaxBytes && pool_alloc(reinterpret_cast<void**>(p), bytes) == CSLAM_OK && *p) return CSLAM_OK;
  *p = nullptr;
  CSLAM_CUDA(cudaMalloc(reinterpret_cast<void**>(p), bytes));
  return CSLAM_OK;
}

template <typename T>
inline void dev_free(T*& p) {
  if (p && !pool_free(p)) cudaFree(p);
  p = nullptr;
}

// Grow-only device buffer.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return CSLAM_OK;
    dev_free(p);
    cap = 0;
    CSLAM_TRY(dev_alloc(&p, n));
    cap = n;
    return CSLAM_OK;
  }
  void release() { dev_free(p); cap = 0; }
};

// ---- device helpers ------------------------------------------------------

// Monotone map float -> uint32 (larger float => larger key).
__device__ __forceinline__ uint32_t f32_to_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_f32(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace cslam
