// Device-resident hash index of the candidate graph: packed 64-bit edge key -> slot of the
// columnar candidate table (incremental graph maintenance with O(1)
// hashed insertion / removal).
//
// Reference: `candidate_edges` is a Python dict keyed by the 4-tuple
// (robot0, keyframe0, robot1, keyframe1) (cslam/algebraic_connectivity_maximization.py:58,150,174);
// add_match (:559-572), remove_candidate_edges (:178-190, a scan of ALL candidates per removed
// edge) and candidate_edges_to_fixed (:192-203) go through it one edge at a time.  Here a batch
// of keys is resolved by one kernel: open addressing with linear probing in HBM, 64-bit keys
// claimed with atomicCAS, tombstones for removals, rebuilt when half full.
#include <vector>

#include "common.cuh"

namespace cslam {
namespace {

constexpr uint64_t kEmpty = ~0ull;
constexpr uint64_t kTomb = ~0ull - 1;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {   // splitmix64 finaliser
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

__global__ void k_km_fill(uint64_t* keys, int64_t cap) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < cap) keys[i] = kEmpty;
}

// values_out[t] = value of keys[t], or -1
__global__ void k_km_lookup(const uint64_t* __restrict__ tkeys, const int32_t* __restrict__ tvals,
                            uint64_t mask, const uint64_t* __restrict__ keys, int64_t n,
                            int32_t* __restrict__ out) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t key = keys[t];
  uint64_t s = mix64(key) & mask;
  int32_t v = -1;
  for (;;) {
    const uint64_t k = tkeys[s];
    if (k == key) { v = tvals[s]; break; }
    if (k == kEmpty) break;
    s = (s + 1) & mask;
  }
  out[t] = v;
}

// insert or overwrite; the keys of one call are distinct.  added[0] counts new keys.
__global__ void k_km_insert(uint64_t* __restrict__ tkeys, int32_t* __restrict__ tvals, uint64_t mask,
                            const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                            int64_t n, unsigned long long* added) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t key = keys[t];
  uint64_t s = mix64(key) & mask;
  for (;;) {
    uint64_t k = tkeys[s];
    if (k == kEmpty) {
      k = atomicCAS(reinterpret_cast<unsigned long long*>(tkeys + s), kEmpty, key);
      if (k == kEmpty) {
        atomicAdd(added, 1ull);
        k = key;
      }
    }
    if (k == key) { tvals[s] = vals[t]; return; }
    s = (s + 1) & mask;   // occupied by another key or a tombstone
  }
}

// erase; out[t] (nullable) = the erased value or -1.  removed[0] counts erased keys.
__global__ void k_km_erase(uint64_t* __restrict__ tkeys, const int32_t* __restrict__ tvals, uint64_t mask,
                           const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ out,
                           unsigned long long* removed) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t key = keys[t];
  uint64_t s = mix64(key) & mask;
  int32_t v = -1;
  for (;;) {
    const uint64_t k = tkeys[s];
    if (k == key) {
      // duplicates of a key inside one call: only the thread that swaps it out counts
      if (atomicCAS(reinterpret_cast<unsigned long long*>(tkeys + s), key, kTomb) == key) {
        v = tvals[s];
        atomicAdd(removed, 1ull);
      }
      break;
    }
    if (k == kEmpty) break;
    s = (s + 1) & mask;
  }
  if (out) out[t] = v;
}

// re-insert every live entry of the old table into the new one
__global__ void k_km_rehash(const uint64_t* __restrict__ okeys, const int32_t* __restrict__ ovals,
                            int64_t ocap, uint64_t* __restrict__ tkeys, int32_t* __restrict__ tvals,
                            uint64_t mask) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= ocap) return;
  const uint64_t key = okeys[i];
  if (key == kEmpty || key == kTomb) return;
  uint64_t s = mix64(key) & mask;
  for (;;) {
    if (atomicCAS(reinterpret_cast<unsigned long long*>(tkeys + s), kEmpty, key) == kEmpty) {
      tvals[s] = ovals[i];
      return;
    }
    s = (s + 1) & mask;
  }
}

}  // namespace
}  // namespace cslam

using namespace cslam;

struct cslam_keymap {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint64_t* d_keys = nullptr;
  int32_t* d_vals = nullptr;
  int64_t cap = 0;        // power of two
  int64_t live = 0;       // keys present
  int64_t used = 0;       // live + tombstones (slots that stop no probe)
  // staging
  uint64_t* d_in_keys = nullptr;
  int32_t* d_in_vals = nullptr;
  unsigned long long* d_count = nullptr;
  int64_t in_cap = 0;
  void* h_pin = nullptr;   // pinned: [in_cap] u64 keys + [in_cap] i32 values + 1 counter
  size_t h_pin_bytes = 0;
};

namespace cslam {
namespace {

int km_alloc_table(cslam_keymap* h, int64_t cap, uint64_t** keys, int32_t** vals) {
  CSLAM_TRY(dev_alloc(keys, static_cast<size_t>(cap)));
  CSLAM_TRY(dev_alloc(vals, static_cast<size_t>(cap)));
  k_km_fill<<<static_cast<unsigned int>((cap + 255) / 256), 256, 0, h->stream>>>(*keys, cap);
  CSLAM_LAUNCH_CHECK();
  return CSLAM_OK;
}

// room for `extra` more keys at a load factor (tombstones included) of at most 1/2
int km_reserve(cslam_keymap* h, int64_t extra) {
  if (2 * (h->used + extra) <= h->cap) return CSLAM_OK;
  int64_t ncap = h->cap > 0 ? h->cap : 1024;
  while (2 * (h->live + extra) > ncap / 2) ncap *= 2;   // rebuilt at a quarter full at most
  uint64_t* nk = nullptr;
  int32_t* nv = nullptr;
  CSLAM_TRY(km_alloc_table(h, ncap, &nk, &nv));
  if (h->cap > 0) {
    k_km_rehash<<<static_cast<unsigned int>((h->cap + 255) / 256), 256, 0, h->stream>>>(
        h->d_keys, h->d_vals, h->cap, nk, nv, static_cast<uint64_t>(ncap - 1));
    CSLAM_LAUNCH_CHECK();
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
  }
  dev_free(h->d_keys);
  dev_free(h->d_vals);
  h->d_keys = nk;
  h->d_vals = nv;
  h->cap = ncap;
  h->used = h->live;
  return CSLAM_OK;
}

int km_stage(cslam_keymap* h, int64_t n) {
  if (n <= h->in_cap) return CSLAM_OK;
  CSLAM_CUDA(cudaStreamSynchronize(h->stream));
  int64_t c = h->in_cap > 0 ? h->in_cap : 4096;
  while (c < n) c *= 2;
  dev_free(h->d_in_keys);
  dev_free(h->d_in_vals);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  h->h_pin = nullptr;
  CSLAM_TRY(dev_alloc(&h->d_in_keys, static_cast<size_t>(c)));
  CSLAM_TRY(dev_alloc(&h->d_in_vals, static_cast<size_t>(c)));
  h->h_pin_bytes = static_cast<size_t>(c) * (sizeof(uint64_t) + sizeof(int32_t)) + 64;
  CSLAM_CUDA(cudaMallocHost(&h->h_pin, h->h_pin_bytes));
  h->in_cap = c;
  return CSLAM_OK;
}

uint64_t* pin_keys(cslam_keymap* h) { return static_cast<uint64_t*>(h->h_pin); }
int32_t* pin_vals(cslam_keymap* h) {
  return reinterpret_cast<int32_t*>(static_cast<char*>(h->h_pin) + static_cast<size_t>(h->in_cap) * sizeof(uint64_t));
}

}  // namespace
}  // namespace cslam

extern "C" {

int cslam_keymap_create(int64_t capacity_hint, int device, cslam_keymap_t** out) {
  CSLAM_REQUIRE(out, "keymap_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("keymap_create: no CUDA device available (this library has no CPU fallback)");
    return CSLAM_ERR_CUDA;
  }
  CSLAM_REQUIRE(device >= 0 && device < ndev, "keymap_create: device %d out of range", device);
  DeviceGuard g(device);
  cslam_keymap* h = new cslam_keymap();
  h->device = device;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&h->d_count), sizeof(unsigned long long)) != cudaSuccess) {
    set_error("keymap_create: resource allocation failed");
    cslam_keymap_destroy(h);
    return CSLAM_ERR_CUDA;
  }
  const int st = km_reserve(h, capacity_hint > 0 ? capacity_hint : 1);
  if (st != CSLAM_OK) {
    cslam_keymap_destroy(h);
    return st;
  }
  *out = h;
  return CSLAM_OK;
}

int cslam_keymap_destroy(cslam_keymap_t* h) {
  if (!h) return CSLAM_OK;
  DeviceGuard g(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  dev_free(h->d_keys);
  dev_free(h->d_vals);
  dev_free(h->d_in_keys);
  dev_free(h->d_in_vals);
  dev_free(h->d_count);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CSLAM_OK;
}

int64_t cslam_keymap_size(cslam_keymap_t* h) { return h ? h->live : 0; }

int cslam_keymap_lookup(cslam_keymap_t* h, const uint64_t* keys, int64_t n, int32_t* values_out) {
  CSLAM_REQUIRE(h && (n == 0 || (keys && values_out)) && n >= 0, "keymap_lookup: bad arguments");
  if (n == 0) return CSLAM_OK;
  DeviceGuard g(h->device);
  CSLAM_TRY(km_stage(h, n));
  memcpy(pin_keys(h), keys, static_cast<size_t>(n) * sizeof(uint64_t));
  CSLAM_CUDA(cudaMemcpyAsync(h->d_in_keys, pin_keys(h), n * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
  k_km_lookup<<<static_cast<unsigned int>((n + 255) / 256), 256, 0, h->stream>>>(
      h->d_keys, h->d_vals, static_cast<uint64_t>(h->cap - 1), h->d_in_keys, n, h->d_in_vals);
  CSLAM_LAUNCH_CHECK();
  CSLAM_CUDA(cudaMemcpyAsync(pin_vals(h), h->d_in_vals, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CSLAM_CUDA(cudaStreamSynchronize(h->stream));
  memcpy(values_out, pin_vals(h), static_cast<size_t>(n) * sizeof(int32_t));
  return CSLAM_OK;
}

int cslam_keymap_insert(cslam_keymap_t* h, const uint64_t* keys, const int32_t* values, int64_t n) {
  CSLAM_REQUIRE(h && (n == 0 || (keys && values)) && n >= 0, "keymap_insert: bad arguments");
  if (n == 0) return CSLAM_OK;
  for (int64_t t = 0; t < n; ++t)
    CSLAM_REQUIRE(keys[t] < kTomb, "keymap_insert: key %lld is reserved", static_cast<long long>(t));
  DeviceGuard g(h->device);
  CSLAM_TRY(km_reserve(h, n));
  CSLAM_TRY(km_stage(h, n));
  memcpy(pin_keys(h), keys, static_cast<size_t>(n) * sizeof(uint64_t));
  memcpy(pin_vals(h), values, static_cast<size_t>(n) * sizeof(int32_t));
  CSLAM_CUDA(cudaMemcpyAsync(h->d_in_keys, pin_keys(h), n * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
  CSLAM_CUDA(cudaMemcpyAsync(h->d_in_vals, pin_vals(h), n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  CSLAM_CUDA(cudaMemsetAsync(h->d_count, 0, sizeof(unsigned long long), h->stream));
  k_km_insert<<<static_cast<unsigned int>((n + 255) / 256), 256, 0, h->stream>>>(
      h->d_keys, h->d_vals, static_cast<uint64_t>(h->cap - 1), h->d_in_keys, h->d_in_vals, n, h->d_count);
  CSLAM_LAUNCH_CHECK();
  unsigned long long added = 0;
  CSLAM_CUDA(cudaMemcpyAsync(&added, h->d_count, sizeof(added), cudaMemcpyDeviceToHost, h->stream));
  CSLAM_CUDA(cudaStreamSynchronize(h->stream));
  h->live += static_cast<int64_t>(added);
  h->used += static_cast<int64_t>(added);
  return CSLAM_OK;
}

int cslam_keymap_erase(cslam_keymap_t* h, const uint64_t* keys, int64_t n, int32_t* values_out) {
  CSLAM_REQUIRE(h && (n == 0 || keys) && n >= 0, "keymap_erase: bad arguments");
  if (n == 0) return CSLAM_OK;
  DeviceGuard g(h->device);
  CSLAM_TRY(km_stage(h, n));
  memcpy(pin_keys(h), keys, static_cast<size_t>(n) * sizeof(uint64_t));
  CSLAM_CUDA(cudaMemcpyAsync(h->d_in_keys, pin_keys(h), n * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
  CSLAM_CUDA(cudaMemsetAsync(h->d_count, 0, sizeof(unsigned long long), h->stream));
  k_km_erase<<<static_cast<unsigned int>((n + 255) / 256), 256, 0, h->stream>>>(
      h->d_keys, h->d_vals, static_cast<uint64_t>(h->cap - 1), h->d_in_keys, n,
      values_out ? h->d_in_vals : nullptr, h->d_count);
  CSLAM_LAUNCH_CHECK();
  unsigned long long removed = 0;
  CSLAM_CUDA(cudaMemcpyAsync(&removed, h->d_count, sizeof(removed), cudaMemcpyDeviceToHost, h->stream));
  if (values_out)
    CSLAM_CUDA(cudaMemcpyAsync(pin_vals(h), h->d_in_vals, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CSLAM_CUDA(cudaStreamSynchronize(h->stream));
  if (values_out) memcpy(values_out, pin_vals(h), static_cast<size_t>(n) * sizeof(int32_t));
  h->live -= static_cast<int64_t>(removed);
  return CSLAM_OK;
}

}  // extern "C"
