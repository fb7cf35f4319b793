// Library-level entry points: version, last-error string, device probe,
// launch counter.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace cslam {

static thread_local char t_err[1024] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

// ---- pooled device allocations (common.cuh dev_alloc / dev_free) ---------------------------------
namespace {
constexpr int kMaxDevices = 64;
struct PoolState {
  std::mutex mu;
  cudaStream_t stream[kMaxDevices] = {};      // one internal non-blocking stream per device
  int state[kMaxDevices] = {};                // 0 untried, 1 ready, -1 unavailable
  std::unordered_map<void*, int> owner;       // pointers handed out by the pool -> device
};
PoolState& pool_state() {
  static PoolState* ps = new PoolState();     // (leaked on purpose: frees may arrive during interpreter shutdown)
  return *ps;
}
bool pool_enabled() {
  static const bool on = !(getenv("CSLAM_DEV_POOL") && atoi(getenv("CSLAM_DEV_POOL")) == 0);
  return on;
}
// the current device's internal stream, or nullptr when the device has no memory pools
cudaStream_t pool_stream(PoolState& ps, int* dev_out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
    cudaGetLastError();
    return nullptr;
  }
  *dev_out = dev;
  if (ps.state[dev] == 0) {
    ps.state[dev] = -1;
    int supported = 0;
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev) == cudaSuccess && supported &&
        cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess &&
        cudaStreamCreateWithFlags(&ps.stream[dev], cudaStreamNonBlocking) == cudaSuccess) {
      uint64_t never = UINT64_MAX;            // keep freed blocks in the pool instead of returning them to the OS
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &never);
      ps.state[dev] = 1;
    }
    cudaGetLastError();
  }
  return ps.state[dev] == 1 ? ps.stream[dev] : nullptr;
}
}  // namespace

int pool_alloc(void** p, size_t bytes) {
  *p = nullptr;
  if (!pool_enabled()) return CSLAM_ERR_CUDA;
  PoolState& ps = pool_state();
  std::lock_guard<std::mutex> lock(ps.mu);
  int dev = 0;
  cudaStream_t s = pool_stream(ps, &dev);
  if (!s) return CSLAM_ERR_CUDA;
  void* q = nullptr;
  if (cudaMallocAsync(&q, bytes, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
    cudaGetLastError();                       // (out of pool memory, ...: the caller falls back to cudaMalloc)
    return CSLAM_ERR_CUDA;
  }
  ps.owner[q] = dev;
  *p = q;
  return CSLAM_OK;
}

bool pool_free(void* p) {
  PoolState& ps = pool_state();
  std::lock_guard<std::mutex> lock(ps.mu);
  auto it = ps.owner.find(p);
  if (it == ps.owner.end()) return false;
  const int dev = it->second;
  ps.owner.erase(it);
  int cur = -1;
  cudaGetDevice(&cur);
  if (cur != dev) cudaSetDevice(dev);
  cudaDeviceSynchronize();                    // what cudaFree does implicitly: nothing in flight may still use p
  if (cudaFreeAsync(p, ps.stream[dev]) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(p);
  }
  if (cur >= 0 && cur != dev) cudaSetDevice(cur);
  return true;
}

}  // namespace cslam

extern "C" {

const char* cslam_version(void) { return "cslam_b200 0.1.0 (sm_100a)"; }

const char* cslam_last_error(void) { return cslam::t_err; }

int cslam_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int64_t cslam_launch_count(void) { return cslam::g_launches.load(); }

}  // extern "C"
