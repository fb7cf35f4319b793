// Multi-robot matching round (cslam_b200/swarm.py): what happens to the all-gathered per-shard
// top-k before anything goes back to the host.
//
// Reference: every robot turns the best match of a keyframe in every OTHER robot's pool into a
// candidate edge when its similarity reaches the threshold
// (cslam/loop_closure_sparse_matching.py:45-53, 62-72), and keeps, of the top
// `nb_best_matches` local matches of its own keyframe, those that were in the pool before the
// keyframe was added (:74-92 with global_descriptor_loop_closure_detection.py:157-160).  Both
// filters run here on the device so that a round ships a handful of hits instead of the whole
// [R, R*B, k] exchange buffers.
#include "common.cuh"

namespace cslam {
namespace {

// items t = (q * B + b) * R + g  (query robot, keyframe, pool robot): the order in which the
// reference meets them.  One block; ordered compaction by ballot + running offset.
__global__ void __launch_bounds__(1024)
k_swarm_hits(int R, int B, int kx, const int64_t* __restrict__ g_kf /*[R][R*B][kx]*/,
             const double* __restrict__ g_sims, const int64_t* __restrict__ all_ids /*[R][B]*/,
             double thr, double* __restrict__ out /*[1 + 5 * cap]*/, int cap) {
  __shared__ int sh_warp[32];
  __shared__ int sh_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) sh_base = 0;
  __syncthreads();
  const int total = R * B * R;
  for (int t0 = 0; t0 < total; t0 += 1024) {
    const int t = t0 + tid;
    bool hit = false;
    int q = 0, b = 0, g = 0;
    int64_t kf1 = -1;
    double s = 0.0;
    if (t < total) {
      g = t % R;
      b = (t / R) % B;
      q = t / (R * B);
      const size_t src = (static_cast<size_t>(g) * R * B + static_cast<size_t>(q) * B + b) * kx;
      kf1 = g_kf[src];
      s = g_sims[src];
      hit = g != q && kf1 >= 0 && s >= thr;     // NaN (empty pool) compares false
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) sh_warp[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
      int v = sh_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      sh_warp[lane] = v;
    }
    __syncthreads();
    const int base = sh_base + (warp > 0 ? sh_warp[warp - 1] : 0) + __popc(m & ((1u << lane) - 1u));
    if (hit && base < cap) {
      double* o = out + 1 + static_cast<size_t>(base) * 5;
      o[0] = q;
      o[1] = static_cast<double>(all_ids[q * B + b]);
      o[2] = g;
      o[3] = static_cast<double>(kf1);
      o[4] = s;
    }
    __syncthreads();
    if (tid == 0) sh_base += sh_warp[31];
    __syncthreads();
  }
  if (tid == 0) out[0] = sh_base;
}

// own keyframe b keeps, in rank order, the matches whose pool row was there before b was
// appended (row < rows_before + b); at most k_keep of them.  One warp per keyframe.
__global__ void k_swarm_intra(int B, int k_search, int k_keep, int64_t rows_before,
                              const int64_t* __restrict__ idx /*[B][k_search] pool rows*/,
                              const int64_t* __restrict__ kf /*[B][k_search]*/,
                              const double* __restrict__ sims,
                              double* __restrict__ out /*[B][1 + 2 * k_keep]*/) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  double* o = out + static_cast<size_t>(b) * (1 + 2 * k_keep);
  int kept = 0;
  for (int j0 = 0; j0 < k_search && kept < k_keep; j0 += 32) {
    const int j = j0 + lane;
    bool keep = false;
    if (j < k_search) {
      const int64_t r = idx[static_cast<size_t>(b) * k_search + j];
      keep = r >= 0 && r < rows_before + b;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int pos = kept + __popc(m & ((1u << lane) - 1u));
    if (keep && pos < k_keep) {
      o[1 + pos] = static_cast<double>(kf[static_cast<size_t>(b) * k_search + j]);
      o[1 + k_keep + pos] = sims[static_cast<size_t>(b) * k_search + j];
    }
    kept += __popc(m);
  }
  if (lane == 0) o[0] = kept < k_keep ? kept : k_keep;
}

}  // namespace
}  // namespace cslam

using namespace cslam;

extern "C" {

int cslam_swarm_hits(int R, int B, int kx, const int64_t* d_g_kf, const double* d_g_sims,
                     const int64_t* d_all_ids, double threshold, double* d_out, int cap, void* stream) {
  CSLAM_REQUIRE(R >= 1 && B >= 1 && kx >= 1 && cap >= 0 && d_g_kf && d_g_sims && d_all_ids && d_out,
                "swarm_hits: bad arguments");
  k_swarm_hits<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(R, B, kx, d_g_kf, d_g_sims, d_all_ids,
                                                               threshold, d_out, cap);
  CSLAM_LAUNCH_CHECK();
  return CSLAM_OK;
}

int cslam_swarm_intra(int B, int k_search, int k_keep, int64_t rows_before, const int64_t* d_idx,
                      const int64_t* d_kf, const double* d_sims, double* d_out, void* stream) {
  CSLAM_REQUIRE(B >= 1 && k_search >= 1 && k_keep >= 1 && d_idx && d_kf && d_sims && d_out,
                "swarm_intra: bad arguments");
  k_swarm_intra<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(B, k_search, k_keep, rows_before,
                                                                           d_idx, d_kf, d_sims, d_out);
  CSLAM_LAUNCH_CHECK();
  return CSLAM_OK;
}

}  // extern "C"
