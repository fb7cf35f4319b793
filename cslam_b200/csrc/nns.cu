// Cosine nearest-neighbour matching over a growable descriptor pool resident in
// HBM.  B200-native replacement for cslam/nns_matching.py:6-76
// (NearestNeighborsMatching): the arithmetic the reference does with a Python
// loop over scipy.spatial.distance.cosine runs here as
//
//   append   : k_nns_append          fp32 row store + float32 |x|^2 + fp16 unit-norm shadow
//   prepare  : k_nns_prep_queries    float64 query copy, |q|^2, fp16 unit-norm query tile
//   candidates: k_nns_coarse_tc      tcgen05 GEMM with fused threshold filter (nns_coarse_tc.cu)
//              k_nns_coarse_exact    fp64 SIMT scan with the same interface (validation /
//                                    escalation path; still on the GPU)
//   threshold: k_nns_tau_select      k-th largest sampled chunk maximum - 2 eps per query
//   result   : k_nns_select          keep {coarse >= c_k - 2 eps} (c_k = k-th largest coarse)
//              k_nns_rerank          exact float64 cosine of the kept float32 rows, exactly
//                                    as the reference scores them
//              k_nns_finalize        sort (similarity desc, row id desc), emit top-k
//
// Exactness argument (spelled out above k_nns_select and in DESIGN.md): with
// |coarse - exact| <= eps for every row (fp16 rounding of unit vectors + fp32
// accumulation), the exact top-k is contained in {coarse >= c_k - 2 eps}, and the
// candidate threshold tau is chosen <= c_k - 2 eps, so the kept set provably contains the
// exact top-k; only capacity limits can fail, and those re-run the query through the
// exact fp64 scan kernel.
#include <cuda_fp16.h>
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "nns_internal.cuh"

namespace cslam {
namespace {

constexpr int kRerankMax = 2048;   // largest re-rank set handled in shared memory
constexpr int kMaxK = 1024;        // largest k accepted by search
constexpr int kSelThreads = 512;
constexpr int kStageRows = 4096;   // host->device staging granularity for add_host

// ---------------------------------------------------------------------------
// k_nns_append: one warp per appended row.
//   data[row, :]   = x                         (float32, the reference's `.data`)
//   vv[row]        = sum_i x_i^2 in float32    (np.dot(v, v) on a float32 row,
//                                               scipy correlation(): vv)
//   shadow[row, :] = fp16(x / sqrt(vv)), zero padded to dim_pad
// ---------------------------------------------------------------------------
__global__ void k_nns_append(const float* __restrict__ src, int64_t count, int dim, int dim_pad,
                             float* __restrict__ data, float* __restrict__ vv,
                             __half* __restrict__ shadow, int64_t row_base) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (w >= count) return;
  const float* x = src + w * dim;
  const int64_t row = row_base + w;
  float acc = 0.f;
  for (int i = lane; i < dim; i += 32) {
    const float v = x[i];
    acc = fmaf(v, v, acc);
  }
  const float s = warp_sum(acc);
  const float inv = s > 0.f ? 1.0f / sqrtf(s) : 0.f;
  float* drow = data + row * dim;
  __half* srow = shadow + row * dim_pad;
  for (int i = lane; i < dim_pad; i += 32) {
    if (i < dim) {
      const float v = x[i];
      drow[i] = v;
      srow[i] = __float2half_rn(v * inv);
    } else {
      srow[i] = __float2half_rn(0.f);
    }
  }
  if (lane == 0) vv[row] = s;
}

// ---------------------------------------------------------------------------
// k_nns_prep_queries: one warp per query row of the padded tile buffer.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void k_nns_prep_queries(const T* __restrict__ q, int nq, int nq_pad, int dim,
                                   int dim_pad, double* __restrict__ q64, double* __restrict__ uu,
                                   __half* __restrict__ qh) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= nq_pad) return;
  __half* hrow = qh + static_cast<size_t>(w) * dim_pad;
  if (w >= nq) {
    for (int i = lane; i < dim_pad; i += 32) hrow[i] = __float2half_rn(0.f);
    return;
  }
  const T* x = q + static_cast<size_t>(w) * dim;
  double acc = 0.0;
  for (int i = lane; i < dim; i += 32) {
    const double v = static_cast<double>(x[i]);
    acc = fma(v, v, acc);
  }
  const double s = warp_sum(acc);
  const double inv = s > 0.0 ? 1.0 / sqrt(s) : 0.0;
  double* drow = q64 + static_cast<size_t>(w) * dim;
  for (int i = lane; i < dim_pad; i += 32) {
    if (i < dim) {
      const double v = static_cast<double>(x[i]);
      drow[i] = v;
      hrow[i] = __float2half_rn(static_cast<float>(v * inv));
    } else {
      hrow[i] = __float2half_rn(0.f);
    }
  }
  if (lane == 0) uu[w] = s;
}

// ---------------------------------------------------------------------------
// k_nns_coarse_exact: same contract as k_nns_coarse_tc but scores in float64
// on the SIMT pipes.  One warp per pool row, looping over the query tile.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_nns_coarse_exact(const float* __restrict__ data, const float* __restrict__ vv, int dim,
                   const double* __restrict__ q64, const double* __restrict__ uu,
                   CoarseParams prm) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
    const int row0 = static_cast<int>(static_cast<int64_t>(tile) * prm.tile_stride * kCoarseBN);
    for (int r = warp; r < kCoarseBN; r += warps_per_block) {
      const int row = row0 + r;
      if (row >= prm.n_rows) break;
      const float* x = data + static_cast<size_t>(row) * dim;
      const double vvd = static_cast<double>(vv[row]);
      for (int q = 0; q < prm.nq; ++q) {
        const double* qv = q64 + static_cast<size_t>(prm.q_row0 + q) * dim;
        double acc = 0.0;
        for (int i = lane; i < dim; i += 32) acc = fma(qv[i], static_cast<double>(x[i]), acc);
        const double uv = warp_sum(acc);
        if (lane == 0) {
          const float s = static_cast<float>(uv / sqrt(uu[prm.q_row0 + q] * vvd));
          if (prm.mode == 1) {
            atomicMax(prm.smax + static_cast<size_t>(q) * prm.smax_stride +
                          tile * (kCoarseBN / kChunk) + r / kChunk,
                      f32_to_key(s));
            continue;
          }
          const float tau = prm.tau ? prm.tau[q] : -INFINITY;
          if (s >= tau) {
            unsigned int* qc = prm.cnt + static_cast<size_t>(q) * (prm.nsub + 1);
            uint2* qcand = prm.cand + static_cast<size_t>(q) * cand_slots(prm.nsub);
            const uint2 e = make_uint2(__float_as_uint(s), static_cast<uint32_t>(row));
            const int sub = blockIdx.x * 2 + (r >= kCoarseBN / 2 ? 1 : 0);
            const unsigned int pos = atomicAdd(qc + sub, 1u);
            if (pos < static_cast<unsigned int>(kSegCap)) {
              qcand[static_cast<size_t>(sub) * kSegCap + pos] = e;
            } else {
              const unsigned int opos = atomicAdd(qc + prm.nsub, 1u);
              if (opos < static_cast<unsigned int>(kOvfCap))
                qcand[static_cast<size_t>(prm.nsub) * kSegCap + opos] = e;
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Block-wide radix select over keys held in shared memory: value of the `kth` largest
// (1-based) of s_key[0..n).  4 passes of 8 bits, MSB first.  s_hist: 256 counters,
// s_ctl: 2 words.  All threads of the block must call it.
// ---------------------------------------------------------------------------
__device__ uint32_t block_kth_largest_smem(const uint32_t* s_key, int n, int kth,
                                           uint32_t* s_hist, uint32_t* s_ctl) {
  uint32_t prefix = 0, mask = 0;
  int remaining = kth;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = s_key[i];
      if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // lane l owns digits [8l, 8l+8); suffix sums locate the digit holding rank `remaining`
      const int l = threadIdx.x;
      uint32_t h[8];
      uint32_t tot = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { h[j] = s_hist[8 * l + j]; tot += h[j]; }
      // above = number of keys in digits owned by higher lanes
      uint32_t incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_down_sync(0xffffffffu, incl, o);
        if (l + o < 32) incl += v;
      }
      const uint32_t above = incl - tot;
      const bool mine = above < static_cast<uint32_t>(remaining) &&
                        static_cast<uint32_t>(remaining) <= above + tot;
      if (mine) {
        uint32_t acc = above;
        int d = 7;
        for (; d > 0; --d) {
          if (acc + h[d] >= static_cast<uint32_t>(remaining)) break;
          acc += h[d];
        }
        s_ctl[0] = static_cast<uint32_t>(8 * l + d);
        s_ctl[1] = static_cast<uint32_t>(remaining) - acc;
      }
    }
    __syncthreads();
    prefix |= s_ctl[0] << shift;
    mask |= 255u << shift;
    remaining = static_cast<int>(s_ctl[1]);
    __syncthreads();
  }
  return prefix;
}

// ---------------------------------------------------------------------------
// k_nns_tau_select: tau[q] = (the `kth` largest of the sampled chunk maxima) - slack.
// Every chunk maximum is a distinct pool row, so at least `kth` pool rows score
// >= that value: the pool's kth largest coarse score is >= tau[q] + slack.
// ---------------------------------------------------------------------------
constexpr int kTauThreads = 256;
constexpr int kTauMax = 4096;  // max sampled chunks per query

__global__ void __launch_bounds__(kTauThreads)
k_nns_tau_select(const uint32_t* __restrict__ smax, int stride, int nchunks, int kth,
                 float slack, float* __restrict__ tau) {
  __shared__ uint32_t s[kTauMax];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_ctl[2];
  const int q = blockIdx.x;
  for (int i = threadIdx.x; i < nchunks; i += blockDim.x)
    s[i] = smax[static_cast<size_t>(q) * stride + i];
  __syncthreads();
  uint32_t kth_key = 0u;
  if (kth <= nchunks) kth_key = block_kth_largest_smem(s, nchunks, kth, s_hist, s_ctl);
  if (threadIdx.x == 0) {
    // round toward -inf: a threshold that is too low only costs candidates
    float t = -INFINITY;
    if (kth <= nchunks) t = __fsub_rd(key_to_f32(kth_key), slack);
    tau[q] = t;
  }
}

// ---------------------------------------------------------------------------
// Candidate-list traversal.  A query's list is `nsub` private sub-segments of
// kSegCap slots (s_cnt[sub] valid entries each) plus the overflow region
// (s_cnt[nsub] entries).  All kSelThreads threads call `f(valid, entry)`
// convergently so that f may use warp collectives.
// ---------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ void for_each_candidate(const uint2* __restrict__ c,
                                                   const uint32_t* s_cnt, int nsub, F f) {
  constexpr int kPerSweep = kSelThreads / kSegCap;
  const int off = threadIdx.x & (kSegCap - 1);
  for (int base = 0; base < nsub; base += kPerSweep) {
    const int sub = base + threadIdx.x / kSegCap;
    const bool valid = sub < nsub && off < static_cast<int>(s_cnt[sub]);
    uint2 e = make_uint2(0u, 0u);
    if (valid) e = c[static_cast<size_t>(sub) * kSegCap + off];
    f(valid, e);
  }
  const int n_ovf = static_cast<int>(s_cnt[nsub]);
  const uint2* o = c + static_cast<size_t>(nsub) * kSegCap;
  for (int base = 0; base < n_ovf; base += kSelThreads) {
    const int i = base + threadIdx.x;
    const bool valid = i < n_ovf;
    uint2 e = make_uint2(0u, 0u);
    if (valid) e = o[i];
    f(valid, e);
  }
}

// Loads the per-sub-segment counts of query q into shared memory (clamped to capacity).
// s_tot[0] = stored candidates, s_tot[1] = 1 if the overflow region dropped entries.
__device__ int load_counts(const unsigned int* __restrict__ cnt, int nsub, uint32_t* s_cnt,
                           int* s_tot) {
  if (threadIdx.x == 0) {
    s_tot[0] = 0;
    s_tot[1] = 0;
  }
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i <= nsub; i += blockDim.x) {
    const unsigned int raw = cnt[i];
    const unsigned int lim = (i == nsub) ? kOvfCap : kSegCap;
    const unsigned int v = raw < lim ? raw : lim;
    s_cnt[i] = v;
    local += static_cast<int>(v);
    if (i == nsub && raw > lim) s_tot[1] = 1;
  }
  if (local) atomicAdd(&s_tot[0], local);
  __syncthreads();
  return s_tot[0];
}

// Block-wide radix select over a candidate list in global memory: key of the `kth`
// largest (1-based) score.  4 passes of 8 bits, MSB first; histogram updates are
// aggregated per warp with match_any.  Only used for lists too long for shared memory.
__device__ uint32_t block_kth_largest_key(const uint2* __restrict__ c, const uint32_t* s_cnt,
                                          int nsub, int kth, uint32_t* s_hist, uint32_t* s_ctl) {
  uint32_t prefix = 0, mask = 0;
  int remaining = kth;
  const int lane = threadIdx.x & 31;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for_each_candidate(c, s_cnt, nsub, [&](bool valid, uint2 e) {
      const uint32_t key = f32_to_key(__uint_as_float(e.x));
      const bool take = valid && ((key & mask) == prefix);
      const unsigned int active = __ballot_sync(0xffffffffu, take);
      if (take) {
        const uint32_t digit = (key >> shift) & 255u;
        const unsigned int peers = __match_any_sync(active, digit);
        if (lane == __ffs(peers) - 1) atomicAdd(&s_hist[digit], __popc(peers));
      }
    });
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0;
      int d = 255;
      for (; d > 0; --d) {
        const int h = static_cast<int>(s_hist[d]);
        if (acc + h >= remaining) break;
        acc += h;
      }
      s_ctl[0] = static_cast<uint32_t>(d);
      s_ctl[1] = static_cast<uint32_t>(remaining - acc);
    }
    __syncthreads();
    prefix |= s_ctl[0] << shift;
    mask |= 255u << shift;
    remaining = static_cast<int>(s_ctl[1]);
    __syncthreads();
  }
  return prefix;
}

// ---------------------------------------------------------------------------
// Result stage, three kernels (so that the exact re-rank spreads over the whole GPU):
//   k_nns_select   (block per query)  sort the candidate list by coarse score, find the
//                  k-th largest coarse score c_k and keep every candidate with
//                  coarse >= c_k - 2*eps  (see the proof below)
//   k_nns_rerank   (warp per kept row) exact float64 cosine from the float32 row, exactly
//                  as the reference scores it (nns_matching.py:58 -> scipy correlation())
//   k_nns_finalize (block per query)  sort by (similarity desc, row id desc), write top-k
//
// Sufficiency of the kept set.  |coarse_j - exact_j| <= eps for every pool row j.  The k
// rows with the largest coarse scores have exact >= c_k - eps, so the exact k-th best
// similarity s_k >= c_k - eps.  A row of the exact top-k has exact >= s_k, hence
// coarse >= s_k - eps >= c_k - 2*eps: it is kept.  The candidate list itself holds every
// row with coarse >= tau, and tau <= c_k - 2*eps by construction (k_nns_tau_select), so no
// qualifying row is missing from the list.  The only ways to fail are capacity limits
// (list overflow, kept set larger than kRerankMax); those raise flag 2 and the query is
// re-run through the exact fp64 scan.
// ---------------------------------------------------------------------------
constexpr int kMaxSub = 2048;   // upper bound on sub-segments per query
constexpr int kListMax = 8192;  // candidate lists up to this length are sorted in smem

struct SelectParams {
  const uint2* cand;
  const unsigned int* cnt;
  int nsub;
  const float* tau;   // nullptr: candidate list is the whole pool
  int n_rows;
  int k;
  float eps;
  int* sel_rows;      // [nq, kRerankMax]
  int* sel_cnt;       // [nq]
  int* flags;         // [nq]: 0 ok, 2 needs exact re-run
};

constexpr size_t kSelSmemBytes =
    sizeof(uint32_t) * kListMax + sizeof(int) * kListMax + sizeof(uint32_t) * (kMaxSub + 1);

__global__ void __launch_bounds__(kSelThreads)
k_nns_select(SelectParams p) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  uint32_t* s_key = reinterpret_cast<uint32_t*>(sel_smem);
  int* s_row = reinterpret_cast<int*>(s_key + kListMax);
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_row + kListMax);  // [kMaxSub + 1]
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_ctl[2];
  __shared__ int s_tot[2];
  __shared__ int s_count;

  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int nsub = p.nsub;
  const uint2* c = p.cand + static_cast<size_t>(q) * cand_slots(nsub);
  const int kk = min(p.k, p.n_rows);
  int* out_rows = p.sel_rows + static_cast<size_t>(q) * kRerankMax;

  const int n_c = load_counts(p.cnt + static_cast<size_t>(q) * (nsub + 1), nsub, s_cnt, s_tot);
  if (s_tot[1] != 0 || n_c < kk) {
    if (tid == 0) { p.flags[q] = 2; p.sel_cnt[q] = 0; }  // dropped candidates: exact re-run
    return;
  }
  if (kk == 0) {
    if (tid == 0) { p.flags[q] = 0; p.sel_cnt[q] = 0; }
    return;
  }
  const double tau_q = p.tau ? static_cast<double>(p.tau[q]) : -INFINITY;
  const bool in_smem = n_c <= kListMax;
  float c_k;  // k-th largest coarse score
  if (in_smem) {
    // Compact the sparse sub-segments into shared memory.  Thread t owns sub-segments
    // [4t, 4t+4) (+ a slice of the overflow region); offsets come from a block-wide
    // exclusive scan of the counts so the copy loops are independent loads.
    constexpr int kPerThread = (kMaxSub + kSelThreads - 1) / kSelThreads;  // 4
    uint32_t my_cnt[kPerThread];
    uint32_t my_sum = 0;
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const int sub = tid * kPerThread + j;
      my_cnt[j] = sub < nsub ? s_cnt[sub] : 0u;
      my_sum += my_cnt[j];
    }
    uint32_t incl = my_sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_hist[tid >> 5] = incl;  // warp totals (s_hist is free here)
    __syncthreads();
    if (tid < 32) {
      const int nw = kSelThreads / 32;
      uint32_t w = tid < nw ? s_hist[tid] : 0u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      s_hist[32 + tid] = w;  // inclusive scan of warp totals
    }
    __syncthreads();
    uint32_t pos = incl - my_sum + ((tid >> 5) > 0 ? s_hist[32 + (tid >> 5) - 1] : 0u);
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const uint2* seg = c + static_cast<size_t>(tid * kPerThread + j) * kSegCap;
      for (uint32_t e = 0; e < my_cnt[j]; ++e) {
        const uint2 v = seg[e];
        s_key[pos] = f32_to_key(__uint_as_float(v.x));
        s_row[pos] = static_cast<int>(v.y);
        ++pos;
      }
    }
    {
      const int n_ovf = static_cast<int>(s_cnt[nsub]);
      const int base = n_c - n_ovf;  // overflow entries go last
      const uint2* o = c + static_cast<size_t>(nsub) * kSegCap;
      for (int i = tid; i < n_ovf; i += blockDim.x) {
        const uint2 v = o[i];
        s_key[base + i] = f32_to_key(__uint_as_float(v.x));
        s_row[base + i] = static_cast<int>(v.y);
      }
    }
    __syncthreads();
    c_k = key_to_f32(block_kth_largest_smem(s_key, n_c, kk, s_hist, s_ctl));
  } else {
    c_k = key_to_f32(block_kth_largest_key(c, s_cnt, nsub, kk, s_hist, s_ctl));
  }
  // keep {coarse >= c_k - 2 eps}; round the threshold DOWN to float (superset)
  const double t2 = static_cast<double>(c_k) - 2.0 * static_cast<double>(p.eps);
  float thr = static_cast<float>(t2);
  if (static_cast<double>(thr) > t2) thr = nextafterf(thr, -INFINITY);
  if (static_cast<double>(thr) < tau_q && n_c < p.n_rows) {
    // rows below the candidate threshold could qualify (cannot happen when tau was chosen
    // by k_nns_tau_select for this k; guards caller-supplied thresholds)
    if (tid == 0) { p.flags[q] = 2; p.sel_cnt[q] = 0; }
    return;
  }
  int count;
  if (in_smem) {
    const uint32_t thr_key = f32_to_key(thr);
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int i = tid; i < n_c; i += blockDim.x) {
      if (s_key[i] >= thr_key) {
        const int pos = atomicAdd(&s_count, 1);
        if (pos < kRerankMax) out_rows[pos] = s_row[i];
      }
    }
    __syncthreads();
    count = s_count;
  } else {
    if (tid == 0) s_count = 0;
    __syncthreads();
    for_each_candidate(c, s_cnt, nsub, [&](bool valid, uint2 e) {
      if (valid && __uint_as_float(e.x) >= thr) {
        const int pos = atomicAdd(&s_count, 1);
        if (pos < kRerankMax) out_rows[pos] = static_cast<int>(e.y);
      }
    });
    __syncthreads();
    count = s_count;
  }
  if (tid == 0) {
    const bool bad = count > kRerankMax || count < kk;
    p.flags[q] = bad ? 2 : 0;
    p.sel_cnt[q] = bad ? 0 : count;
  }
}

// One warp per (query, kept row).  grid.x covers kRerankMax rows, grid.y = queries.
__global__ void __launch_bounds__(256)
k_nns_rerank(const int* __restrict__ sel_rows, const int* __restrict__ sel_cnt,
             const float* __restrict__ data, const float* __restrict__ vv, int dim,
             const double* __restrict__ q64, const double* __restrict__ uu,
             double* __restrict__ sel_sims) {
  const int q = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int cnt = sel_cnt[q];
  const double* qv = q64 + static_cast<size_t>(q) * dim;
  // grid-stride over the kept rows (typically ~2k of them): a small fixed grid.x instead of
  // kRerankMax / 8 blocks per query, most of which would exit at once
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < cnt;
       i += gridDim.x * (blockDim.x >> 5)) {
    const int row = sel_rows[static_cast<size_t>(q) * kRerankMax + i];
    const float* x = data + static_cast<size_t>(row) * dim;
    double acc = 0.0;
    for (int d = lane; d < dim; d += 32) acc = fma(qv[d], static_cast<double>(x[d]), acc);
    const double uv = warp_sum(acc);
    if (lane == 0) {
      double dist = 1.0 - uv / sqrt(uu[q] * static_cast<double>(vv[row]));
      dist = fmin(fmax(dist, 0.0), 2.0);
      double sim = 1.0 - dist;
      if (!(sim == sim)) sim = -INFINITY;  // NaN (zero-norm row) ranks last
      sel_sims[static_cast<size_t>(q) * kRerankMax + i] = sim;
    }
  }
}

__device__ __forceinline__ bool ranks_before(double sa, int ia, double sb, int ib) {
  // descending similarity, ties by descending row id
  return (sa > sb) || (sa == sb && ia > ib);
}

__global__ void __launch_bounds__(kSelThreads)
k_nns_finalize(const int* __restrict__ sel_rows, const int* __restrict__ sel_cnt,
               const double* __restrict__ sel_sims, const int* __restrict__ flags, int n_rows,
               int k, int32_t* __restrict__ out_idx, double* __restrict__ out_sims) {
  __shared__ double s_sim[kRerankMax];
  __shared__ int s_idx[kRerankMax];
  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  int32_t* oi = out_idx + static_cast<size_t>(q) * k;
  double* os = out_sims + static_cast<size_t>(q) * k;
  const int kk = min(k, n_rows);
  const int count = sel_cnt[q];
  if (flags[q] != 0 || count < kk) {
    for (int j = tid; j < k; j += blockDim.x) { oi[j] = -1; os[j] = nan(""); }
    return;
  }
  int P = 1;
  while (P < count) P <<= 1;
  for (int i = tid; i < P; i += blockDim.x) {
    const bool v = i < count;
    s_sim[i] = v ? sel_sims[static_cast<size_t>(q) * kRerankMax + i] : -INFINITY;
    s_idx[i] = v ? sel_rows[static_cast<size_t>(q) * kRerankMax + i] : -1;
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int st = size >> 1; st > 0; st >>= 1) {
      for (int t = tid; t < (P >> 1); t += blockDim.x) {
        const int lo = ((t & ~(st - 1)) << 1) | (t & (st - 1));  // st is a power of two
        const int hi = lo + st;
        const bool desc_block = ((lo & size) == 0);
        const double sa = s_sim[lo], sb = s_sim[hi];
        const int ia = s_idx[lo], ib = s_idx[hi];
        const bool a_first = ranks_before(sa, ia, sb, ib);
        if (desc_block ? !a_first : a_first) {
          s_sim[lo] = sb; s_sim[hi] = sa;
          s_idx[lo] = ib; s_idx[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < k; j += blockDim.x) {
    oi[j] = j < kk ? s_idx[j] : -1;
    os[j] = j < kk ? s_sim[j] : nan("");
  }
}

__global__ void k_fill_empty(int32_t* idx, double* sims, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    idx[i] = -1;
    sims[i] = nan("");
  }
}

}  // namespace
}  // namespace cslam

// ---------------------------------------------------------------------------
// Handle
// ---------------------------------------------------------------------------
using namespace cslam;

struct cslam_nns {
  int device = 0;
  int dim = 0;
  int dim_pad = 0;
  int num_sms = 148;
  int64_t n = 0;
  int64_t cap = 0;
  float* d_data = nullptr;
  __half* d_shadow = nullptr;
  float* d_vv = nullptr;
  TensorMapBlob tmap_p[3];   // pool tile maps with box rows 256, 128, 64 (cluster size 1, 2, 4)
  int max_clusters[3] = {0, 0, 0};   // co-resident clusters of size 1, 2, 4
  cudaStream_t stream = nullptr;

  // host -> device staging for add_host
  float* h_stage = nullptr;   // pinned [kStageRows, dim]
  float* d_stage = nullptr;   // [kStageRows, dim]
  int64_t staged = 0;

  // search workspace
  int q_cap = 0;              // rows in the query buffers (multiple of kCoarseBM)
  void* d_qraw = nullptr;     // raw queries for the host variant (f64-sized)
  double* d_q64 = nullptr;
  double* d_uu = nullptr;
  __half* d_qh = nullptr;
  TensorMapBlob tmap_q;
  float* d_tau = nullptr;         // [kCoarseBM]
  int* d_flags = nullptr;         // [kCoarseBM]
  int* d_sel_rows = nullptr;      // [kCoarseBM, kRerankMax] rows kept for the exact re-rank
  int* d_sel_cnt = nullptr;       // [kCoarseBM]
  double* d_sel_sims = nullptr;   // [kCoarseBM, kRerankMax]
  int* h_flags = nullptr;         // pinned [kCoarseBM]
  DevBuf<uint2> cand;             // [kCoarseBM, cand_slots(nsub)]
  unsigned int* d_cnt = nullptr;  // [kCoarseBM, nsub + 1]
  int nsub = 0;
  DevBuf<uint32_t> smax;          // [kCoarseBM, sample chunks] sampled chunk maxima (keys)
  int out_cap_q = 0, out_cap_k = 0;
  int32_t* d_out_idx = nullptr;
  double* d_out_sims = nullptr;
  void* h_q = nullptr;        // pinned query staging
  size_t h_q_bytes = 0;
  int32_t* h_out_idx = nullptr;
  double* h_out_sims = nullptr;
  size_t h_out_elems = 0;

  int mode = 0;
  int sample_rows = 32768;

  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_c0 = nullptr, ev_c1 = nullptr;
  // Cross-stream ordering of everything that touches the pool or the search workspace: the
  // last operation's stream and an event recorded behind it.  An operation on another stream
  // first waits for that event (nns_enter), so appends, growth copies and searches issued on
  // the caller's streams and on the handle's own stream execute in call order.
  cudaEvent_t ev_order = nullptr;
  cudaStream_t last_stream = nullptr;
  bool has_order = false;
  int last_coarse_launches = 0;
  bool timing_valid = false;
};

namespace cslam {
namespace {

// Make stream `s` wait for the handle's previous operation if that ran on another stream.
int nns_enter(cslam_nns* h, cudaStream_t s) {
  if (h->has_order && h->last_stream != s) CSLAM_CUDA(cudaStreamWaitEvent(s, h->ev_order, 0));
  return CSLAM_OK;
}

// Mark the end of an operation issued on `s`.
int nns_leave(cslam_nns* h, cudaStream_t s) {
  CSLAM_CUDA(cudaEventRecord(h->ev_order, s));
  h->last_stream = s;
  h->has_order = true;
  return CSLAM_OK;
}

// Growth copies run on the stream of the append that needs them (`s`, already ordered behind
// the handle's previous operation by nns_enter), so rows appended earlier on that stream are
// complete before they are moved.
int nns_reserve_rows(cslam_nns* h, int64_t need, cudaStream_t s) {
  if (need <= h->cap) return CSLAM_OK;
  int64_t ncap = h->cap > 0 ? h->cap : 1024;
  while (ncap < need) ncap *= 2;
  ncap = (ncap + kCoarseBN - 1) / kCoarseBN * kCoarseBN;
  if (ncap > 0x7fffff00ll) {
    set_error("descriptor pool limited to 2^31 rows");
    return CSLAM_ERR_LIMIT;
  }
  float* nd = nullptr;
  __half* ns = nullptr;
  float* nv = nullptr;
  CSLAM_TRY(dev_alloc(&nd, static_cast<size_t>(ncap) * h->dim));
  CSLAM_TRY(dev_alloc(&ns, static_cast<size_t>(ncap) * h->dim_pad));
  CSLAM_TRY(dev_alloc(&nv, static_cast<size_t>(ncap)));
  CSLAM_CUDA(cudaMemsetAsync(ns, 0, static_cast<size_t>(ncap) * h->dim_pad * sizeof(__half),
                             s));
  if (h->n > 0) {
    CSLAM_CUDA(cudaMemcpyAsync(nd, h->d_data, static_cast<size_t>(h->n) * h->dim * sizeof(float),
                               cudaMemcpyDeviceToDevice, s));
    CSLAM_CUDA(cudaMemcpyAsync(ns, h->d_shadow,
                               static_cast<size_t>(h->n) * h->dim_pad * sizeof(__half),
                               cudaMemcpyDeviceToDevice, s));
    CSLAM_CUDA(cudaMemcpyAsync(nv, h->d_vv, static_cast<size_t>(h->n) * sizeof(float),
                               cudaMemcpyDeviceToDevice, s));
  }
  CSLAM_CUDA(cudaStreamSynchronize(s));
  dev_free(h->d_data);
  dev_free(h->d_shadow);
  dev_free(h->d_vv);
  h->d_data = nd;
  h->d_shadow = ns;
  h->d_vv = nv;
  h->cap = ncap;
  for (int c = 0; c < 3; ++c)
    CSLAM_TRY(make_fp16_rowmajor_tmap(&h->tmap_p[c], h->d_shadow, h->cap, h->dim_pad, kCoarseBN >> c));
  return CSLAM_OK;
}

int nns_append_device(cslam_nns* h, const float* d_rows, int64_t count, cudaStream_t s) {
  if (count <= 0) return CSLAM_OK;
  CSLAM_TRY(nns_reserve_rows(h, h->n + count, s));
  const int threads = 256;
  const int64_t blocks = (count * 32 + threads - 1) / threads;
  k_nns_append<<<static_cast<unsigned int>(blocks), threads, 0, s>>>(
      d_rows, count, h->dim, h->dim_pad, h->d_data, h->d_vv, h->d_shadow, h->n);
  CSLAM_LAUNCH_CHECK();
  h->n += count;
  return CSLAM_OK;
}

int nns_flush(cslam_nns* h) {
  if (h->staged == 0) return CSLAM_OK;
  CSLAM_TRY(nns_enter(h, h->stream));
  CSLAM_CUDA(cudaMemcpyAsync(h->d_stage, h->h_stage,
                             static_cast<size_t>(h->staged) * h->dim * sizeof(float),
                             cudaMemcpyHostToDevice, h->stream));
  CSLAM_TRY(nns_append_device(h, h->d_stage, h->staged, h->stream));
  // the pinned staging buffer is reused by the next add: wait for the copy
  CSLAM_CUDA(cudaStreamSynchronize(h->stream));
  CSLAM_TRY(nns_leave(h, h->stream));
  h->staged = 0;
  return CSLAM_OK;
}

int nns_reserve_queries(cslam_nns* h, int nq, int k) {
  const int need = (nq + kCoarseBM - 1) / kCoarseBM * kCoarseBM;
  if (need > h->q_cap) {
    dev_free(h->d_q64);
    dev_free(h->d_uu);
    dev_free(h->d_qh);
    if (h->d_qraw) { cudaFree(h->d_qraw); h->d_qraw = nullptr; }
    CSLAM_TRY(dev_alloc(&h->d_q64, static_cast<size_t>(need) * h->dim));
    CSLAM_TRY(dev_alloc(&h->d_uu, static_cast<size_t>(need)));
    CSLAM_TRY(dev_alloc(&h->d_qh, static_cast<size_t>(need) * h->dim_pad));
    CSLAM_CUDA(cudaMalloc(&h->d_qraw, static_cast<size_t>(need) * h->dim * sizeof(double)));
    h->q_cap = need;
    CSLAM_TRY(make_fp16_rowmajor_tmap(&h->tmap_q, h->d_qh, need, h->dim_pad, kCoarseBM));
  }
  if (!h->d_tau) {
    CSLAM_TRY(dev_alloc(&h->d_tau, kGroupMax));
    CSLAM_TRY(dev_alloc(&h->d_flags, kGroupMax));
    CSLAM_TRY(dev_alloc(&h->d_sel_rows, static_cast<size_t>(kGroupMax) * kRerankMax));
    CSLAM_TRY(dev_alloc(&h->d_sel_cnt, kGroupMax));
    CSLAM_TRY(dev_alloc(&h->d_sel_sims, static_cast<size_t>(kGroupMax) * kRerankMax));
    CSLAM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->h_flags), 2 * kGroupMax * sizeof(int)));
  }
  (void)k;
  return CSLAM_OK;
}

// One query GROUP (nqt <= kGroupMax queries starting at q0; <= kCoarseBM for the exact scan)
// through candidate generation + re-rank.  Groups wider than one tile run the tensor-core
// pass as clusters of 2 or 4 CTAs that share every pool tile (TMA multicast), so the pool is
// swept once per group.  `exact` selects the fp64 scan instead of the tensor cores.  Results
// land in out_idx/out_sims (device, [*, k], already offset).
int nns_run_tile(cslam_nns* h, int q0, int nqt, int k, bool exact, int32_t* out_idx,
                 double* out_sims, cudaStream_t s, bool time_it) {
  const int n_rows = static_cast<int>(h->n);
  const int num_tiles = (n_rows + kCoarseBN - 1) / kCoarseBN;
  const bool exhaustive = n_rows <= h->sample_rows;
  const int sample_tiles = h->sample_rows / kCoarseBN;
  const int q_tiles = (nqt + kCoarseBM - 1) / kCoarseBM;
  int cl = exact ? 0 : (q_tiles <= 1 ? 0 : (q_tiles <= 2 ? 1 : 2));   // log2(cluster size)
  const int C = 1 << cl;
  // two query tiles: SM-pair kernel (tcgen05 cta_group::2), one pool sweep per 256 queries
  static const bool pair_enabled = !(getenv("CSLAM_NNS_PAIR") && atoi(getenv("CSLAM_NNS_PAIR")) == 0);
  const bool use_pair = !exact && C == 2 && pair_enabled;
  CSLAM_REQUIRE(nqt <= (exact ? kCoarseBM : kGroupMax), "nns: query group of %d too wide", nqt);
  if (!exact && h->max_clusters[cl] == 0) {
    if (C == 1) h->max_clusters[0] = h->num_sms;
    else CSLAM_TRY(coarse_tc_max_clusters(C, &h->max_clusters[cl]));
    CSLAM_REQUIRE(h->max_clusters[cl] > 0, "nns: no co-resident cluster of %d CTAs", C);
    if (getenv("CSLAM_NNS_DEBUG"))
      fprintf(stderr, "[cslam nns] cluster size %d: %d co-resident clusters (%d SMs)\n", C,
              h->max_clusters[cl], h->num_sms);
  }
  const int resident = exact ? h->num_sms : h->max_clusters[cl];   // clusters (C = 1: CTAs)
  // dump passes (tau = -inf, small pools) give every cluster exactly one tile, i.e. each
  // sub-segment exactly kSegCap entries; the filtered pass runs as many clusters as fit.
  const int max_grid = std::max(h->num_sms, sample_tiles);
  const int nsub_needed = 2 * max_grid;
  if (nsub_needed > h->nsub) {
    CSLAM_REQUIRE(nsub_needed <= kMaxSub, "nns: %d sub-segments exceed the limit %d", nsub_needed,
                  kMaxSub);
    CSLAM_CUDA(cudaStreamSynchronize(s));
    CSLAM_TRY(h->cand.reserve(cand_slots(nsub_needed) * kGroupMax));
    dev_free(h->d_cnt);
    CSLAM_TRY(dev_alloc(&h->d_cnt, static_cast<size_t>(kGroupMax) * (nsub_needed + 1)));
    h->nsub = nsub_needed;
  }
  // Wide groups sample more rows for the threshold (131072 instead of 32768): the candidate
  // lists shrink ~4x (k / sample fraction rows clear tau), which is what the epilogue of the
  // tensor-bound kernels pays for; single-tile searches are HBM-bound and keep the small sample.
  int samp_tiles = sample_tiles;
  if (C > 1 && !exhaustive) {
    const int wide = kTauMax / (kCoarseBN / kChunk);          // 512 tiles = 131072 rows
    if (n_rows / kCoarseBN >= 4 * wide) samp_tiles = wide;
  }
  const int smax_stride = samp_tiles * (kCoarseBN / kChunk);
  CSLAM_REQUIRE(smax_stride <= kTauMax, "nns: sample_rows too large");
  CSLAM_TRY(h->smax.reserve(static_cast<size_t>(kGroupMax) * smax_stride));
  // fp16 rounding of both unit-norm operands (2 * 2^-11), fp32 accumulation over
  // dim_pad terms, fp16 subnormal flush; exact scan only rounds the score to fp32.
  const float eps = exact ? 1.0e-6f
                          : (0.0009765625f + static_cast<float>(h->dim_pad) * 1.1920929e-7f +
                             4.0e-6f);

  CoarseParams cp;
  cp.mode = 0;
  cp.num_kb = h->dim_pad / 64;
  cp.n_rows = n_rows;
  cp.nq = nqt;
  cp.q_row0 = q0;
  cp.cluster = C;
  cp.l2_prefetch = getenv("CSLAM_NNS_PF") ? atoi(getenv("CSLAM_NNS_PF")) : 0;   // measured: no gain, the ring is not latency-bound
  cp.a_resident = (C > 1 && h->dim_pad <= 512 && !getenv("CSLAM_NNS_NO_ARES")) ? 1 : 0;
  cp.cnt = h->d_cnt;
  cp.cand = h->cand.p;
  cp.nsub = h->nsub;
  cp.smax = h->smax.p;
  cp.smax_stride = smax_stride;
  cp.dbg = nullptr;
  static long long* s_dbg = nullptr;
  if (getenv("CSLAM_NNS_DEBUG")) {
    if (!s_dbg) { cudaMalloc(&s_dbg, 8 * sizeof(long long)); }
    cudaMemsetAsync(s_dbg, 0, 8 * sizeof(long long), s);
    cp.dbg = s_dbg;
  }

  auto launch_coarse = [&](int mode, int tiles, int stride, const float* tau, bool timed) -> int {
    cp.mode = mode;
    cp.num_tiles = tiles;
    cp.tile_stride = stride;
    cp.tau = tau;
    if (mode == 0) {
      CSLAM_CUDA(cudaMemsetAsync(
          h->d_cnt, 0, static_cast<size_t>(nqt) * (h->nsub + 1) * sizeof(unsigned int), s));
    } else if (exact) {
      CSLAM_CUDA(cudaMemsetAsync(h->smax.p, 0,
                                 static_cast<size_t>(nqt) * smax_stride * sizeof(uint32_t),
                                 s));
    }
    if (timed) CSLAM_CUDA(cudaEventRecord(h->ev_c0, s));
    // unfiltered passes: one tile per cluster; filtered full pass: persistent, one cluster per
    // co-resident slot
    const int clusters = tau == nullptr ? std::min(tiles, max_grid) : std::min(tiles, resident);
    if (exact) {
      k_nns_coarse_exact<<<clusters, 256, 0, s>>>(h->d_data, h->d_vv, h->dim, h->d_q64, h->d_uu, cp);
      CSLAM_LAUNCH_CHECK();
    } else if (use_pair) {
      CSLAM_TRY(launch_coarse_pair(&h->tmap_q, &h->tmap_p[1], cp, clusters * 2, s));
    } else {
      CSLAM_TRY(launch_coarse_tc(&h->tmap_q, &h->tmap_p[cl], cp, clusters * C, s));
    }
    if (timed) CSLAM_CUDA(cudaEventRecord(h->ev_c1, s));
    h->last_coarse_launches++;
    return CSLAM_OK;
  };

  const float* tau = nullptr;
  if (exhaustive) {
    CSLAM_TRY(launch_coarse(0, num_tiles, 1, nullptr, time_it));
  } else {
    // sample pass over `sample_tiles` full tiles spread evenly over the pool
    const int full_tiles = n_rows / kCoarseBN;
    const int stride = std::max(1, full_tiles / samp_tiles);
    CSLAM_TRY(launch_coarse(1, samp_tiles, stride, nullptr, false));
    // tau = (k-th largest sampled chunk maximum) - 2 eps.  The sampled maxima are distinct
    // pool rows, so the pool's k-th largest coarse score c_k >= that value and therefore
    // tau <= c_k - 2 eps: every row k_nns_select needs is in the candidate list.
    k_nns_tau_select<<<nqt, kTauThreads, 0, s>>>(h->smax.p, smax_stride, smax_stride,
                                                  std::min(k, n_rows), 2.0f * eps, h->d_tau);
    CSLAM_LAUNCH_CHECK();
    tau = h->d_tau;
    CSLAM_TRY(launch_coarse(0, num_tiles, 1, tau, time_it));
  }

  if (cp.dbg) {
    long long hd[8];
    cudaMemcpyAsync(hd, cp.dbg, sizeof(hd), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    fprintf(stderr, "[cslam nns] CTA0 cycles: producer wait-empty %lld | mma wait-full %lld wait-acc %lld | epilogue wait-acc %lld scan %lld | tiles %lld\n",
            hd[0], hd[1], hd[2], hd[3], hd[4], hd[5]);
  }
  SelectParams sp;
  sp.cand = h->cand.p;
  sp.cnt = h->d_cnt;
  sp.nsub = h->nsub;
  sp.tau = tau;
  sp.n_rows = n_rows;
  sp.k = k;
  sp.eps = eps;
  sp.sel_rows = h->d_sel_rows;
  sp.sel_cnt = h->d_sel_cnt;
  sp.flags = h->d_flags;
  CSLAM_CUDA(cudaFuncSetAttribute(k_nns_select, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(kSelSmemBytes)));
  k_nns_select<<<nqt, kSelThreads, kSelSmemBytes, s>>>(sp);
  CSLAM_LAUNCH_CHECK();
  // one warp per kept row; blocks beyond a query's kept count exit immediately
  const int rr_rows = exhaustive ? std::min(kRerankMax, std::max(n_rows, 1)) : kRerankMax;
  dim3 rgrid(std::min((rr_rows + 7) / 8, 16), nqt);
  k_nns_rerank<<<rgrid, 256, 0, s>>>(h->d_sel_rows, h->d_sel_cnt, h->d_data, h->d_vv, h->dim,
                                     h->d_q64 + static_cast<size_t>(q0) * h->dim, h->d_uu + q0,
                                     h->d_sel_sims);
  CSLAM_LAUNCH_CHECK();
  k_nns_finalize<<<nqt, kSelThreads, 0, s>>>(h->d_sel_rows, h->d_sel_cnt, h->d_sel_sims,
                                             h->d_flags, n_rows, k, out_idx, out_sims);
  CSLAM_LAUNCH_CHECK();
  return CSLAM_OK;
}

template <typename T>
int launch_prep(cslam_nns* h, const void* d_q, int nq, int nq_pad, cudaStream_t s) {
  const int threads = 256;
  const int blocks = (nq_pad * 32 + threads - 1) / threads;
  k_nns_prep_queries<T><<<blocks, threads, 0, s>>>(static_cast<const T*>(d_q), nq, nq_pad, h->dim,
                                                   h->dim_pad, h->d_q64, h->d_uu, h->d_qh);
  CSLAM_LAUNCH_CHECK();
  return CSLAM_OK;
}

int nns_search_impl(cslam_nns* h, const void* d_queries, int dtype, int nq, int k,
                    int32_t* d_out_idx, double* d_out_sims, cudaStream_t s, int64_t* out_info) {
  CSLAM_REQUIRE(nq >= 0 && k >= 1 && k <= kMaxK, "search: need nq >= 0 and 1 <= k <= %d (k=%d)",
                kMaxK, k);
  CSLAM_REQUIRE(dtype == CSLAM_DTYPE_F32 || dtype == CSLAM_DTYPE_F64, "search: bad dtype %d",
                dtype);
  const int64_t launches0 = g_launches.load();
  int64_t info[4] = {0, 0, 0, 0};
  h->last_coarse_launches = 0;
  h->timing_valid = false;
  if (nq == 0) {
    if (out_info) memcpy(out_info, info, sizeof(info));
    return CSLAM_OK;
  }
  CSLAM_TRY(nns_flush(h));
  if (h->n == 0) {
    CSLAM_TRY(nns_enter(h, s));
    const size_t tot = static_cast<size_t>(nq) * k;
    k_fill_empty<<<static_cast<unsigned int>((tot + 255) / 256), 256, 0, s>>>(d_out_idx,
                                                                             d_out_sims, tot);
    CSLAM_LAUNCH_CHECK();
    CSLAM_TRY(nns_leave(h, s));
    if (out_info) {
      info[3] = g_launches.load() - launches0;
      memcpy(out_info, info, sizeof(info));
    }
    return CSLAM_OK;
  }
  CSLAM_TRY(nns_enter(h, s));
  CSLAM_TRY(nns_reserve_queries(h, nq, k));
  const int nq_pad = (nq + kCoarseBM - 1) / kCoarseBM * kCoarseBM;
  if (dtype == CSLAM_DTYPE_F32) {
    CSLAM_TRY(launch_prep<float>(h, d_queries, nq, nq_pad, s));
  } else {
    CSLAM_TRY(launch_prep<double>(h, d_queries, nq, nq_pad, s));
  }
  CSLAM_CUDA(cudaEventRecord(h->ev_t0, s));
  const bool exact = (h->mode == 1);
  // widest group one pool sweep serves: 256 with the SM-pair kernel (HBM-bound per sweep), 512 with
  // the cluster-of-4 multicast kernel (CSLAM_NNS_PAIR=0)
  const bool pair_on = !(getenv("CSLAM_NNS_PAIR") && atoi(getenv("CSLAM_NNS_PAIR")) == 0);
  int group = exact ? kCoarseBM : (pair_on ? 2 * kCoarseBM : kGroupMax);
  if (!exact && getenv("CSLAM_NNS_GROUP")) group = std::max(kCoarseBM, std::min(kGroupMax, atoi(getenv("CSLAM_NNS_GROUP"))));
  for (int q0 = 0; q0 < nq; q0 += group) {
    const int nqt = std::min(group, nq - q0);
    int32_t* oi = d_out_idx + static_cast<size_t>(q0) * k;
    double* os = d_out_sims + static_cast<size_t>(q0) * k;
    CSLAM_TRY(nns_run_tile(h, q0, nqt, k, exact, oi, os, s, /*time_it=*/q0 + group >= nq));
    // escalation check: read the per-query flags
    CSLAM_CUDA(cudaMemcpyAsync(h->h_flags, h->d_flags, nqt * sizeof(int), cudaMemcpyDeviceToHost,
                               s));
    CSLAM_CUDA(cudaMemcpyAsync(h->h_flags + kGroupMax, h->d_sel_cnt, nqt * sizeof(int),
                               cudaMemcpyDeviceToHost, s));
    CSLAM_CUDA(cudaStreamSynchronize(s));
    std::vector<int> redo;
    for (int i = 0; i < nqt; ++i) {
      if (h->h_flags[i] == 2) redo.push_back(i);
      else info[0]++;
      info[1] += h->h_flags[kGroupMax + i];
    }
    for (int i : redo) {
      if (exact) {
        set_error("nns search: exact scan could not resolve query %d (more than %d tied rows?)",
                  q0 + i, kRerankMax);
        return CSLAM_ERR_LIMIT;
      }
      CSLAM_TRY(nns_run_tile(h, q0 + i, 1, k, true, oi + static_cast<size_t>(i) * k,
                             os + static_cast<size_t>(i) * k, s, false));
      CSLAM_CUDA(cudaMemcpyAsync(h->h_flags, h->d_flags, sizeof(int), cudaMemcpyDeviceToHost, s));
      CSLAM_CUDA(cudaStreamSynchronize(s));
      if (h->h_flags[0] == 2) {
        set_error("nns search: exact scan could not resolve query %d (more than %d tied rows?)",
                  q0 + i, kRerankMax);
        return CSLAM_ERR_LIMIT;
      }
      info[2]++;
    }
  }
  CSLAM_CUDA(cudaEventRecord(h->ev_t1, s));
  CSLAM_TRY(nns_leave(h, s));
  h->timing_valid = true;
  info[3] = g_launches.load() - launches0;
  if (out_info) memcpy(out_info, info, sizeof(info));
  return CSLAM_OK;
}

}  // namespace
}  // namespace cslam

extern "C" {

int cslam_nns_create(int dim, int device, cslam_nns_t** out) {
  CSLAM_REQUIRE(out != nullptr, "nns_create: out is NULL");
  *out = nullptr;
  CSLAM_REQUIRE(dim > 0 && dim <= 65536, "nns_create: dim must be in [1, 65536] (got %d)", dim);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("nns_create: no CUDA device available (this library has no CPU fallback)");
    return CSLAM_ERR_CUDA;
  }
  CSLAM_REQUIRE(device >= 0 && device < ndev, "nns_create: device %d out of range [0,%d)", device,
                ndev);
  DeviceGuard g(device);
  if (!g.ok) {
    set_error("nns_create: cannot select device %d", device);
    return CSLAM_ERR_CUDA;
  }
  cudaDeviceProp prop;
  CSLAM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("nns_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
              prop.major, prop.minor);
    return CSLAM_ERR_CUDA;
  }
  cslam_nns* h = new cslam_nns();
  h->device = device;
  h->dim = dim;
  h->dim_pad = (dim + 63) / 64 * 64;
  h->num_sms = prop.multiProcessorCount;
  int st = CSLAM_OK;
  do {
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { st = CSLAM_ERR_CUDA; break; }
    if (cudaMallocHost(reinterpret_cast<void**>(&h->h_stage),
                       static_cast<size_t>(kStageRows) * dim * sizeof(float)) != cudaSuccess) { st = CSLAM_ERR_OOM; break; }
    if (cudaMalloc(reinterpret_cast<void**>(&h->d_stage),
                   static_cast<size_t>(kStageRows) * dim * sizeof(float)) != cudaSuccess) { st = CSLAM_ERR_OOM; break; }
    if (cudaEventCreate(&h->ev_t0) != cudaSuccess || cudaEventCreate(&h->ev_t1) != cudaSuccess ||
        cudaEventCreate(&h->ev_c0) != cudaSuccess || cudaEventCreate(&h->ev_c1) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming) != cudaSuccess) { st = CSLAM_ERR_CUDA; break; }
  } while (0);
  if (st != CSLAM_OK) {
    set_error("nns_create: resource allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    cslam_nns_destroy(h);
    return st;
  }
  *out = h;
  return CSLAM_OK;
}

int cslam_nns_destroy(cslam_nns_t* h) {
  if (!h) return CSLAM_OK;
  DeviceGuard g(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->has_order) cudaEventSynchronize(h->ev_order);
  dev_free(h->d_data);
  dev_free(h->d_shadow);
  dev_free(h->d_vv);
  dev_free(h->d_stage);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  dev_free(h->d_q64);
  dev_free(h->d_uu);
  dev_free(h->d_qh);
  if (h->d_qraw) cudaFree(h->d_qraw);
  dev_free(h->d_tau);
  dev_free(h->d_cnt);
  dev_free(h->d_flags);
  dev_free(h->d_sel_rows);
  dev_free(h->d_sel_cnt);
  dev_free(h->d_sel_sims);
  if (h->h_flags) cudaFreeHost(h->h_flags);
  h->cand.release();
  h->smax.release();
  dev_free(h->d_out_idx);
  dev_free(h->d_out_sims);
  if (h->h_q) cudaFreeHost(h->h_q);
  if (h->h_out_idx) cudaFreeHost(h->h_out_idx);
  if (h->h_out_sims) cudaFreeHost(h->h_out_sims);
  if (h->ev_t0) cudaEventDestroy(h->ev_t0);
  if (h->ev_t1) cudaEventDestroy(h->ev_t1);
  if (h->ev_c0) cudaEventDestroy(h->ev_c0);
  if (h->ev_c1) cudaEventDestroy(h->ev_c1);
  if (h->ev_order) cudaEventDestroy(h->ev_order);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CSLAM_OK;
}

int cslam_nns_add_host(cslam_nns_t* h, const void* rows, int dtype, int64_t count) {
  CSLAM_REQUIRE(h && (rows || count == 0) && count >= 0, "nns_add_host: bad arguments");
  CSLAM_REQUIRE(dtype == CSLAM_DTYPE_F32 || dtype == CSLAM_DTYPE_F64, "nns_add_host: bad dtype %d",
                dtype);
  DeviceGuard g(h->device);
  int64_t done = 0;
  while (done < count) {
    const int64_t take = std::min<int64_t>(count - done, kStageRows - h->staged);
    float* dst = h->h_stage + h->staged * h->dim;
    const size_t elems = static_cast<size_t>(take) * h->dim;
    if (dtype == CSLAM_DTYPE_F32) {
      memcpy(dst, static_cast<const float*>(rows) + done * h->dim, elems * sizeof(float));
    } else {
      const double* src = static_cast<const double*>(rows) + done * h->dim;
      for (size_t i = 0; i < elems; ++i) dst[i] = static_cast<float>(src[i]);  // `.data` is float32
    }
    h->staged += take;
    done += take;
    if (h->staged == kStageRows) CSLAM_TRY(nns_flush(h));
  }
  return CSLAM_OK;
}

int cslam_nns_add_device(cslam_nns_t* h, const float* d_rows, int64_t count, void* stream) {
  CSLAM_REQUIRE(h && (d_rows || count == 0) && count >= 0, "nns_add_device: bad arguments");
  DeviceGuard g(h->device);
  CSLAM_TRY(nns_flush(h));
  // NULL is CUDA's legacy default stream (torch's default stream), never a private one
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CSLAM_TRY(nns_enter(h, s));
  CSLAM_TRY(nns_append_device(h, d_rows, count, s));
  return nns_leave(h, s);
}

int64_t cslam_nns_size(cslam_nns_t* h) { return h ? h->n + h->staged : 0; }
int cslam_nns_dim(cslam_nns_t* h) { return h ? h->dim : 0; }

int cslam_nns_read_rows(cslam_nns_t* h, int64_t start, int64_t count, float* out) {
  CSLAM_REQUIRE(h && out && start >= 0 && count >= 0, "nns_read_rows: bad arguments");
  DeviceGuard g(h->device);
  CSLAM_TRY(nns_flush(h));
  CSLAM_REQUIRE(start + count <= h->n, "nns_read_rows: range [%lld,%lld) exceeds size %lld",
                static_cast<long long>(start), static_cast<long long>(start + count),
                static_cast<long long>(h->n));
  if (count == 0) return CSLAM_OK;
  if (h->has_order) CSLAM_CUDA(cudaEventSynchronize(h->ev_order));
  CSLAM_CUDA(cudaMemcpy(out, h->d_data + start * h->dim,
                        static_cast<size_t>(count) * h->dim * sizeof(float),
                        cudaMemcpyDeviceToHost));
  return CSLAM_OK;
}

int cslam_nns_search_device(cslam_nns_t* h, const void* d_queries, int dtype, int nq, int k,
                            int32_t* d_out_idx, double* d_out_sims, void* stream,
                            int64_t* out_info) {
  CSLAM_REQUIRE(h && (nq == 0 || (d_queries && d_out_idx && d_out_sims)),
                "nns_search_device: NULL argument");
  DeviceGuard g(h->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);   // NULL = legacy default stream
  return nns_search_impl(h, d_queries, dtype, nq, k, d_out_idx, d_out_sims, s, out_info);
}

int cslam_nns_search_host(cslam_nns_t* h, const void* queries, int dtype, int nq, int k,
                          int32_t* out_idx, double* out_sims, int64_t* out_info) {
  CSLAM_REQUIRE(h && (nq == 0 || (queries && out_idx && out_sims)),
                "nns_search_host: NULL argument");
  CSLAM_REQUIRE(nq >= 0 && k >= 1 && k <= kMaxK, "search: need nq >= 0 and 1 <= k <= %d (k=%d)",
                kMaxK, k);
  CSLAM_REQUIRE(dtype == CSLAM_DTYPE_F32 || dtype == CSLAM_DTYPE_F64, "search: bad dtype %d",
                dtype);
  if (nq == 0) return nns_search_impl(h, nullptr, dtype, 0, k, nullptr, nullptr, nullptr, out_info);
  DeviceGuard g(h->device);
  cudaStream_t s = h->stream;
  const size_t esz = dtype == CSLAM_DTYPE_F32 ? sizeof(float) : sizeof(double);
  const size_t qbytes = static_cast<size_t>(nq) * h->dim * esz;
  CSLAM_TRY(nns_reserve_queries(h, nq, k));
  if (qbytes > h->h_q_bytes) {
    if (h->h_q) cudaFreeHost(h->h_q);
    h->h_q = nullptr;
    h->h_q_bytes = 0;
    CSLAM_CUDA(cudaMallocHost(&h->h_q, qbytes));
    h->h_q_bytes = qbytes;
  }
  const size_t oel = static_cast<size_t>(nq) * k;
  if (oel > h->h_out_elems) {
    if (h->h_out_idx) cudaFreeHost(h->h_out_idx);
    if (h->h_out_sims) cudaFreeHost(h->h_out_sims);
    h->h_out_idx = nullptr;
    h->h_out_sims = nullptr;
    dev_free(h->d_out_idx);
    dev_free(h->d_out_sims);
    h->h_out_elems = 0;
    CSLAM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->h_out_idx), oel * sizeof(int32_t)));
    CSLAM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->h_out_sims), oel * sizeof(double)));
    CSLAM_TRY(dev_alloc(&h->d_out_idx, oel));
    CSLAM_TRY(dev_alloc(&h->d_out_sims, oel));
    h->h_out_elems = oel;
  }
  memcpy(h->h_q, queries, qbytes);
  CSLAM_CUDA(cudaMemcpyAsync(h->d_qraw, h->h_q, qbytes, cudaMemcpyHostToDevice, s));
  CSLAM_TRY(nns_search_impl(h, h->d_qraw, dtype, nq, k, h->d_out_idx, h->d_out_sims, s, out_info));
  CSLAM_CUDA(cudaMemcpyAsync(h->h_out_idx, h->d_out_idx, oel * sizeof(int32_t),
                             cudaMemcpyDeviceToHost, s));
  CSLAM_CUDA(cudaMemcpyAsync(h->h_out_sims, h->d_out_sims, oel * sizeof(double),
                             cudaMemcpyDeviceToHost, s));
  CSLAM_CUDA(cudaStreamSynchronize(s));
  memcpy(out_idx, h->h_out_idx, oel * sizeof(int32_t));
  memcpy(out_sims, h->h_out_sims, oel * sizeof(double));
  return CSLAM_OK;
}

int cslam_nns_set_mode(cslam_nns_t* h, int mode) {
  CSLAM_REQUIRE(h && (mode == 0 || mode == 1), "nns_set_mode: mode must be 0 or 1");
  h->mode = mode;
  return CSLAM_OK;
}

int cslam_nns_set_sample_rows(cslam_nns_t* h, int sample_rows) {
  CSLAM_REQUIRE(h, "nns_set_sample_rows: NULL handle");
  CSLAM_REQUIRE(sample_rows >= kCoarseBN && sample_rows % kCoarseBN == 0 &&
                    sample_rows / kChunk <= kTauMax,
                "nns_set_sample_rows: must be a multiple of %d in [%d, %d]", kCoarseBN, kCoarseBN,
                kTauMax * kChunk);
  h->sample_rows = sample_rows;
  return CSLAM_OK;
}

int cslam_nns_last_timing(cslam_nns_t* h, float* coarse_ms, int* coarse_launches,
                          float* total_ms) {
  CSLAM_REQUIRE(h, "nns_last_timing: NULL handle");
  CSLAM_REQUIRE(h->timing_valid, "nns_last_timing: no completed search to report");
  DeviceGuard g(h->device);
  CSLAM_CUDA(cudaEventSynchronize(h->ev_t1));
  float c = 0.f, t = 0.f;
  CSLAM_CUDA(cudaEventElapsedTime(&c, h->ev_c0, h->ev_c1));
  CSLAM_CUDA(cudaEventElapsedTime(&t, h->ev_t0, h->ev_t1));
  if (coarse_ms) *coarse_ms = c;
  if (coarse_launches) *coarse_launches = h->last_coarse_launches;
  if (total_ms) *total_ms = t;
  return CSLAM_OK;
}

}  // extern "C"
