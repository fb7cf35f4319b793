// Lidar place recognition: Scan Context matching on the GPU.
//
// Replaces cslam/lidar_pr/scancontext_matching.py:6-104 (ScanContextMatching) and the two
// helpers it calls, cslam/lidar_pr/scancontext_utils.py:78-79 (sc2rk) and :81-113 (distance_sc).
// The reference answers a query in two steps:
//   1. ring key = mean of every ring (row) of the [rings, sectors] descriptor; the
//      `num_candidates` pool entries whose ring keys are nearest (Euclidean, scipy KDTree rebuilt
//      on every call, :61-66);
//   2. for each candidate, the column-shift distance: for every shift s = 1..sectors the mean
//      cosine similarity between column j of the query and column (j - s) mod sectors of the
//      candidate, over the columns that are non-zero in both; distance = 1 - max_s (:69-79).
// Everything is float64, like the reference's numpy arrays.
//
// HBM layout per handle (capacity doubles like the reference's resize, :33-37):
//   cols [cap][sectors][rings]  every descriptor stored COLUMN by column (a column = one sector,
//                               contiguous) -- the unit step 2 works on
//   norm [cap][sectors]         2-norm of every column, computed once on insert
//   rk   [rings][cap]           ring keys, ring-major so that a warp scanning 32 consecutive
//                               entries reads 256 contiguous bytes per ring (step 1 is an
//                               HBM-bound scan: rings*8 B per entry and query batch)
// Kernels:
//   k_sc_prepare   transpose + column norms + ring key (numpy's pairwise summation order)
//   k_sc_knn       per (range of the pool, tile of up to 8 queries): ring keys staged through
//                  shared memory once per tile, one warp per query keeps its nearest entries in
//                  a lane-distributed sorted list (vote + shift insertion)
//   k_sc_knn_merge per query: merge of the per-range lists
//   k_sc_distance  per (query, candidate): the sectors x sectors cosine table in shared memory,
//                  one ordered sum per shift, first maximum
//   k_sc_pick      per query: first candidate with the smallest distance below 1 (:70-79)
#include <math.h>

#include <vector>

#include "common.cuh"

namespace cslam {
namespace {

constexpr int kMaxCand = 16;       // candidates per query kept in registers
constexpr int kKnnThreads = 256;
constexpr int kDistThreads = 256;

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE),
// which np.mean(sc, axis=1) applies to every contiguous row: 8 running sums for up to 128
// elements, recursive halves above.
__device__ double np_pairwise_sum(const double* a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  if (n <= 128) {
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// One CTA per descriptor.  in: [count][rings][sectors] (float32 or float64).
// cols/norm are indexed by (first_row + b); rk by ring * rk_stride + first_row + b.
template <typename T>
__global__ void k_sc_prepare(const T* __restrict__ in, int rings, int sectors, double* cols,
                             double* norm, double* rk, int64_t rk_stride, int64_t first_row) {
  extern __shared__ double s_in[];   // [rings][sectors]
  const int b = blockIdx.x;
  const int cells = rings * sectors;
  const T* src = in + static_cast<int64_t>(b) * cells;
  for (int e = threadIdx.x; e < cells; e += blockDim.x) s_in[e] = static_cast<double>(src[e]);
  __syncthreads();
  const int64_t row = first_row + b;
  for (int j = threadIdx.x; j < sectors; j += blockDim.x) {
    double ss = 0.0;
    double* dst = cols + (row * sectors + j) * rings;
    for (int r = 0; r < rings; ++r) {
      const double v = s_in[r * sectors + j];
      dst[r] = v;
      ss += v * v;
    }
    norm[row * sectors + j] = sqrt(ss);
  }
  for (int r = threadIdx.x; r < rings; r += blockDim.x)
    rk[r * rk_stride + row] = np_pairwise_sum(s_in + r * sectors, sectors) / static_cast<double>(sectors);
}

__device__ __forceinline__ bool closer(double d, int i, double d2, int i2) {
  return d < d2 || (d == d2 && i < i2);
}

// Sorted list of the warp's nearest entries, one per lane: lane p < kMaxCand holds the p-th
// nearest (ld, li); the other lanes hold +inf.  A candidate that beats the current
// `ncand`-th entry is inserted by a vote (its position = number of entries that stay ahead
// of it) and a one-lane shift: ~10 warp instructions, and it happens ~ncand*ln(rows/ncand)
// times per scan, so the scan itself is what the kernel spends its time on.
struct WarpList {
  double ld;
  int li;
  double thr;     // current ncand-th entry
  int thr_i;
  __device__ __forceinline__ void init() {
    ld = INFINITY;
    li = 0x7fffffff;
    thr = INFINITY;
    thr_i = 0x7fffffff;
  }
  // every lane offers one (d, i); i == 0x7fffffff marks "nothing"
  __device__ __forceinline__ void offer(double d, int i, int ncand, int lane) {
    unsigned mask = __ballot_sync(0xffffffffu, closer(d, i, thr, thr_i));
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const double bd = __shfl_sync(0xffffffffu, d, src);
      const int bi = __shfl_sync(0xffffffffu, i, src);
      if (!closer(bd, bi, thr, thr_i)) continue;     // the threshold moved meanwhile
      const int pos = __popc(__ballot_sync(0xffffffffu, closer(ld, li, bd, bi)));
      const double ud = __shfl_up_sync(0xffffffffu, ld, 1);
      const int ui = __shfl_up_sync(0xffffffffu, li, 1);
      if (lane < kMaxCand) {
        if (lane > pos) { ld = ud; li = ui; }
        if (lane == pos) { ld = bd; li = bi; }
      }
      thr = __shfl_sync(0xffffffffu, ld, ncand - 1);
      thr_i = __shfl_sync(0xffffffffu, li, ncand - 1);
    }
  }
};

constexpr int kKnnWarps = kKnnThreads / 32;   // 8
constexpr int kKnnRows = 256;                 // pool entries staged per step

// Step 1.  grid (row ranges, query tiles of up to 8).  The CTA streams its range of the ring-key
// table through shared memory 256 entries at a time ([rings][256] doubles, loaded once and used
// by every query of the tile); warp w scans the staged entries for query (w % qslots), row part
// (w / qslots) -- 8 queries x 1 part, 4 x 2, 2 x 4 or 1 x 8 -- and keeps its nearest entries in a
// WarpList.  part_*: [queries][parts][ncand], parts = ranges * row parts.
// Algorithmic traffic: rings*8 B per pool entry per query TILE (not per query).
template <int RINGS>   // > 0: compile-time ring count (query key in registers); 0: any count
__global__ void __launch_bounds__(kKnnThreads)
k_sc_knn(const double* __restrict__ rk, int64_t rk_stride, int64_t n, int rings_rt,
         const double* __restrict__ qrk, int nq, int ncand, int64_t rows_per_range, int qslots,
         double* part_d, int* part_i) {
  const int rings = RINGS > 0 ? RINGS : rings_rt;
  extern __shared__ double s_rk[];            // [rings][kKnnRows]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rparts = kKnnWarps / qslots;
  const int slot = warp % qslots, rpart = warp / qslots;
  const int q = blockIdx.y * qslots + slot;
  const bool active = q < nq;
  // this warp's query ring key in registers
  double qk[RINGS > 0 ? RINGS : 1];
  if (RINGS > 0) {
#pragma unroll
    for (int r = 0; r < RINGS; ++r) qk[r] = active ? qrk[static_cast<int64_t>(r) * nq + q] : 0.0;
  }
  WarpList wl;
  wl.init();
  const int64_t lo = blockIdx.x * rows_per_range;
  const int64_t hi = min(n, lo + rows_per_range);
  for (int64_t base = lo; base < hi; base += kKnnRows) {
    const int cnt = static_cast<int>(min(static_cast<int64_t>(kKnnRows), hi - base));
    __syncthreads();
    for (int e = threadIdx.x; e < rings * kKnnRows; e += kKnnThreads) {
      const int r = e / kKnnRows, c = e % kKnnRows;
      s_rk[e] = c < cnt ? rk[r * rk_stride + base + c] : 0.0;
    }
    __syncthreads();
    if (!active) continue;
    for (int c = rpart * 32 + lane; c - lane < cnt; c += 32 * rparts) {   // warp-uniform trip count
      double d = 0.0;
      if (RINGS > 0) {
#pragma unroll
        for (int r = 0; r < RINGS; ++r) {
          const double t = s_rk[r * kKnnRows + c] - qk[r];
          d = __dadd_rn(d, __dmul_rn(t, t));     // the reference's order of additions, no fma
        }
      } else {
        for (int r = 0; r < rings; ++r) {
          const double t = s_rk[r * kKnnRows + c] - qrk[static_cast<int64_t>(r) * nq + q];
          d = __dadd_rn(d, __dmul_rn(t, t));
        }
      }
      const bool valid = c < cnt;
      wl.offer(valid ? d : INFINITY, valid ? static_cast<int>(base + c) : 0x7fffffff, ncand, lane);
    }
  }
  if (active && lane < ncand) {
    const int64_t parts = static_cast<int64_t>(gridDim.x) * rparts;
    const int64_t slot_out = (static_cast<int64_t>(q) * parts + blockIdx.x * rparts + rpart) * ncand + lane;
    part_d[slot_out] = wl.li == 0x7fffffff ? INFINITY : wl.ld;
    part_i[slot_out] = wl.li == 0x7fffffff ? -1 : wl.li;
  }
}

// grid (queries).  Merge `parts` sorted lists of `ncand` into cand_*: [queries][ncand].
__global__ void __launch_bounds__(kKnnThreads)
k_sc_knn_merge(const double* __restrict__ part_d, const int* __restrict__ part_i, int parts,
               int ncand, double* cand_d, int* cand_i) {
  __shared__ double s_d[kKnnWarps][kMaxCand];
  __shared__ int s_i[kKnnWarps][kMaxCand];
  const int q = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpList wl;
  wl.init();
  const int64_t base = static_cast<int64_t>(q) * parts * ncand;
  const int total = parts * ncand;
  for (int e = warp * 32 + lane; e - lane < total; e += kKnnThreads) {
    const bool valid = e < total && part_i[base + e] >= 0;
    wl.offer(valid ? part_d[base + e] : INFINITY, valid ? part_i[base + e] : 0x7fffffff, ncand, lane);
  }
  if (lane < kMaxCand) { s_d[warp][lane] = wl.ld; s_i[warp][lane] = wl.li; }
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < kKnnWarps; ++w) {
      const bool has = lane < kMaxCand;
      wl.offer(has ? s_d[w][lane] : INFINITY, has ? s_i[w][lane] : 0x7fffffff, ncand, lane);
    }
    if (lane < ncand) {
      cand_d[q * ncand + lane] = wl.li == 0x7fffffff ? INFINITY : wl.ld;
      cand_i[q * ncand + lane] = wl.li == 0x7fffffff ? -1 : wl.li;
    }
  }
}

// grid (ncand, queries).  distance_sc(candidate, query) of scancontext_utils.py:81-113.
__global__ void __launch_bounds__(kDistThreads)
k_sc_distance(const double* __restrict__ cols, const double* __restrict__ norm,
              const double* __restrict__ qcols, const double* __restrict__ qnorm,
              const int* __restrict__ cand_i, int ncand, int rings, int sectors, double* out_dist,
              int* out_yaw) {
  extern __shared__ double sm[];
  double* a = sm;                           // candidate columns [sectors][rings]
  double* b = a + sectors * rings;          // query columns
  double* na = b + sectors * rings;         // [sectors]; < 0 marks an all-zero column
  double* nb = na + sectors;
  double* cs = nb + sectors;                // [shift][column] cosine, NaN = column not engaged
  double* sim = cs + sectors * sectors;     // [shift]
  const int c = blockIdx.x, q = blockIdx.y;
  const int row = cand_i[q * ncand + c];
  if (row < 0) {                            // fewer pool entries than candidates
    if (threadIdx.x == 0) { out_dist[q * ncand + c] = 1.0; out_yaw[q * ncand + c] = 1; }
    return;
  }
  const int cells = sectors * rings;
  const double* ga = cols + static_cast<int64_t>(row) * cells;
  const double* gb = qcols + static_cast<int64_t>(q) * cells;
  for (int e = threadIdx.x; e < cells; e += kDistThreads) { a[e] = ga[e]; b[e] = gb[e]; }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * sectors; j += kDistThreads) {
    const bool second = j >= sectors;
    const int col = second ? j - sectors : j;
    const double* v = (second ? b : a) + col * rings;
    bool any = false;
    for (int r = 0; r < rings; ++r) any = any || (v[r] != 0.0);   // ~np.any(col): skip
    const double nv = second ? qnorm[static_cast<int64_t>(q) * sectors + col]
                             : norm[static_cast<int64_t>(row) * sectors + col];
    (second ? nb : na)[col] = any ? nv : -1.0;
  }
  __syncthreads();
  // after s one-column rolls, column j of the candidate is its original column (j - s) mod sectors:
  // the cosine of candidate column ja and query column j belongs to shift s = (j - ja) mod sectors
  // (0 -> sectors).  2 x 2 register tiles over the (ja, j) plane: 40 shared loads per 80 FMAs.
  {
    const int half = (sectors + 1) / 2;
    for (int tile = threadIdx.x; tile < half * half; tile += kDistThreads) {
      const int ja0 = 2 * (tile / half), j0 = 2 * (tile % half);
      const bool a1 = ja0 + 1 < sectors, b1 = j0 + 1 < sectors;
      const double* x0 = a + ja0 * rings;
      const double* x1 = a + (a1 ? ja0 + 1 : ja0) * rings;
      const double* y0 = b + j0 * rings;
      const double* y1 = b + (b1 ? j0 + 1 : j0) * rings;
      double d00 = 0.0, d01 = 0.0, d10 = 0.0, d11 = 0.0;
      for (int r = 0; r < rings; ++r) {
        const double xa = x0[r], xb = x1[r], ya = y0[r], yb = y1[r];
        d00 += xa * ya;
        d01 += xa * yb;
        d10 += xb * ya;
        d11 += xb * yb;
      }
      auto put = [&](int ja, int j, double dot) {
        int sh = j - ja;
        sh += sh <= 0 ? sectors : 0;                       // 1 .. sectors
        const bool live = na[ja] >= 0.0 && nb[j] >= 0.0;
        cs[(sh - 1) * sectors + j] = live ? dot / (na[ja] * nb[j]) : nan("");
      };
      put(ja0, j0, d00);
      if (b1) put(ja0, j0 + 1, d01);
      if (a1) put(ja0 + 1, j0, d10);
      if (a1 && b1) put(ja0 + 1, j0 + 1, d11);
    }
  }
  __syncthreads();
  for (int s = threadIdx.x; s < sectors; s += kDistThreads) {
    double sum = 0.0;
    int engaged = 0;
    for (int j = 0; j < sectors; ++j) {
      const double v = cs[s * sectors + j];
      if (v == v) { sum += v; ++engaged; }
    }
    sim[s] = engaged ? sum / engaged : 0.0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = 0;
    for (int s = 1; s < sectors; ++s)
      if (sim[s] > sim[best]) best = s;     // np.argmax: first maximum
    out_dist[q * ncand + c] = 1.0 - sim[best];
    out_yaw[q * ncand + c] = best + 1;
  }
}

// One thread per query: the loop of scancontext_matching.py:69-79 and the fallback of :81-87.
__global__ void k_sc_pick(const double* __restrict__ dist, const int* __restrict__ yaw,
                          const int* __restrict__ cand_i, int nq, int ncand, int* out_row,
                          double* out_sim, int* out_yaw) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  double nn = 1.0;
  int row = -1, y = 0;
  for (int c = 0; c < ncand; ++c) {
    const double d = dist[q * ncand + c];
    if (cand_i[q * ncand + c] >= 0 && d < nn) { nn = d; row = cand_i[q * ncand + c]; y = yaw[q * ncand + c]; }
  }
  out_row[q] = row;                          // -1: no candidate closer than 1 -> item 0, similarity 0
  out_sim[q] = row < 0 ? 0.0 : 1.0 - nn;
  out_yaw[q] = row < 0 ? 0 : y;
}

}  // namespace
}  // namespace cslam

using namespace cslam;

struct cslam_sc {
  int rings = 0, sectors = 0, ncand = 0, device = 0;
  cudaStream_t stream = nullptr;
  int64_t n = 0, cap = 0;
  double* cols = nullptr;
  double* norm = nullptr;
  double* rk = nullptr;
  DevBuf<unsigned char> staging;
  DevBuf<double> qcols, qnorm, qrk, part_d, cand_d, dist, sim;
  DevBuf<int> part_i, cand_i, yaw, best_row, best_yaw;
  float knn_ms = 0.f, dist_ms = 0.f;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

namespace {

int sc_grow(cslam_sc* h, int64_t need) {
  if (need <= h->cap) return CSLAM_OK;
  int64_t cap = h->cap ? h->cap : 1000;       // the reference starts at 1000 and doubles (:18-19,33-37)
  while (cap < need) cap *= 2;
  const int64_t cells = static_cast<int64_t>(h->rings) * h->sectors;
  double *cols = nullptr, *norm = nullptr, *rk = nullptr;
  CSLAM_TRY(dev_alloc(&cols, cap * cells));
  CSLAM_TRY(dev_alloc(&norm, cap * h->sectors));
  CSLAM_TRY(dev_alloc(&rk, cap * h->rings));
  if (h->n > 0) {
    CSLAM_CUDA(cudaMemcpyAsync(cols, h->cols, h->n * cells * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CSLAM_CUDA(cudaMemcpyAsync(norm, h->norm, h->n * h->sectors * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CSLAM_CUDA(cudaMemcpy2DAsync(rk, cap * sizeof(double), h->rk, h->cap * sizeof(double),
                                 h->n * sizeof(double), h->rings, cudaMemcpyDeviceToDevice, h->stream));
  }
  CSLAM_CUDA(cudaStreamSynchronize(h->stream));
  dev_free(h->cols);
  dev_free(h->norm);
  dev_free(h->rk);
  h->cols = cols;
  h->norm = norm;
  h->rk = rk;
  h->cap = cap;
  return CSLAM_OK;
}

// Upload `count` descriptors and run k_sc_prepare into the given destination.
int sc_prepare(cslam_sc* h, const void* host, int dtype, int64_t count, double* cols, double* norm,
               double* rk, int64_t rk_stride, int64_t first_row) {
  const int64_t cells = static_cast<int64_t>(h->rings) * h->sectors;
  const size_t esz = dtype == CSLAM_DTYPE_F64 ? 8 : 4;
  CSLAM_TRY(h->staging.reserve(count * cells * esz));
  CSLAM_CUDA(cudaMemcpyAsync(h->staging.p, host, count * cells * esz, cudaMemcpyHostToDevice, h->stream));
  const size_t smem = cells * sizeof(double);
  if (dtype == CSLAM_DTYPE_F64)
    k_sc_prepare<double><<<static_cast<unsigned>(count), 64, smem, h->stream>>>(
        reinterpret_cast<const double*>(h->staging.p), h->rings, h->sectors, cols, norm, rk, rk_stride, first_row);
  else
    k_sc_prepare<float><<<static_cast<unsigned>(count), 64, smem, h->stream>>>(
        reinterpret_cast<const float*>(h->staging.p), h->rings, h->sectors, cols, norm, rk, rk_stride, first_row);
  CSLAM_LAUNCH_CHECK();
  return CSLAM_OK;
}

size_t dist_smem(int rings, int sectors) {
  return (2ull * sectors * rings + 2ull * sectors + 1ull * sectors * sectors + sectors) * sizeof(double);
}

}  // namespace

extern "C" {

int cslam_sc_create(int rings, int sectors, int num_candidates, int device, cslam_sc_t** out) {
  CSLAM_REQUIRE(out != nullptr, "cslam_sc_create: out is NULL");
  *out = nullptr;
  CSLAM_REQUIRE(rings > 0 && sectors > 0 && rings <= 96 && sectors <= 1024,
                "cslam_sc_create: shape [%d, %d] out of range", rings, sectors);
  CSLAM_REQUIRE(num_candidates >= 1 && num_candidates <= kMaxCand,
                "cslam_sc_create: num_candidates must be 1..%d", kMaxCand);
  if (cslam_device_count() <= 0) {
    set_error("cslam_sc_create: no CUDA device (there is no CPU fallback)");
    return CSLAM_ERR_CUDA;
  }
  if (dist_smem(rings, sectors) > 200 * 1024 || static_cast<size_t>(rings) * sectors * 8 > 200 * 1024) {
    set_error("cslam_sc_create: shape [%d, %d] does not fit the shared-memory tiles", rings, sectors);
    return CSLAM_ERR_LIMIT;
  }
  DeviceGuard g(device);
  if (!g.ok) { set_error("cslam_sc_create: cannot select device %d", device); return CSLAM_ERR_CUDA; }
  cslam_sc* h = new cslam_sc();
  h->rings = rings;
  h->sectors = sectors;
  h->ncand = num_candidates;
  h->device = device;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    set_error("cslam_sc_create: cudaStreamCreate failed");
    return CSLAM_ERR_CUDA;
  }
  for (auto& e : h->ev) cudaEventCreate(&e);
  cudaFuncSetAttribute(k_sc_distance, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       static_cast<int>(dist_smem(rings, sectors)));
  cudaFuncSetAttribute(k_sc_prepare<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, rings * sectors * 8);
  cudaFuncSetAttribute(k_sc_prepare<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, rings * sectors * 8);
  cudaFuncSetAttribute(k_sc_knn<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, rings * kKnnRows * 8);
  cudaFuncSetAttribute(k_sc_knn<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, rings * kKnnRows * 8);
  int s = sc_grow(h, 1000);
  if (s != CSLAM_OK) { cslam_sc_destroy(h); return s; }
  *out = h;
  return CSLAM_OK;
}

int cslam_sc_destroy(cslam_sc_t* h) {
  if (!h) return CSLAM_OK;
  DeviceGuard g(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  dev_free(h->cols);
  dev_free(h->norm);
  dev_free(h->rk);
  h->staging.release(); h->qcols.release(); h->qnorm.release(); h->qrk.release();
  h->part_d.release(); h->cand_d.release(); h->dist.release(); h->sim.release();
  h->part_i.release(); h->cand_i.release(); h->yaw.release(); h->best_row.release(); h->best_yaw.release();
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CSLAM_OK;
}

int64_t cslam_sc_size(cslam_sc_t* h) { return h ? h->n : -1; }
int64_t cslam_sc_capacity(cslam_sc_t* h) { return h ? h->cap : -1; }

int cslam_sc_add_host(cslam_sc_t* h, const void* descriptors, int dtype, int64_t count) {
  CSLAM_REQUIRE(h != nullptr, "cslam_sc_add_host: NULL handle");
  CSLAM_REQUIRE(dtype == CSLAM_DTYPE_F32 || dtype == CSLAM_DTYPE_F64, "cslam_sc_add_host: bad dtype %d", dtype);
  CSLAM_REQUIRE(count >= 0 && (count == 0 || descriptors != nullptr), "cslam_sc_add_host: bad arguments");
  if (count == 0) return CSLAM_OK;
  CSLAM_REQUIRE(h->n + count < (1ll << 31), "cslam_sc_add_host: pool limited to 2^31 entries");
  DeviceGuard g(h->device);
  CSLAM_TRY(sc_grow(h, h->n + count));
  CSLAM_TRY(sc_prepare(h, descriptors, dtype, count, h->cols, h->norm, h->rk, h->cap, h->n));
  CSLAM_CUDA(cudaStreamSynchronize(h->stream));   // the caller may reuse its buffer
  h->n += count;
  return CSLAM_OK;
}

int cslam_sc_read(cslam_sc_t* h, int64_t start, int64_t count, double* out_scancontexts,
                  double* out_ringkeys) {
  CSLAM_REQUIRE(h != nullptr, "cslam_sc_read: NULL handle");
  CSLAM_REQUIRE(start >= 0 && count >= 0 && start + count <= h->n, "cslam_sc_read: rows [%lld, %lld) out of range",
                static_cast<long long>(start), static_cast<long long>(start + count));
  if (count == 0) return CSLAM_OK;
  DeviceGuard g(h->device);
  const int64_t cells = static_cast<int64_t>(h->rings) * h->sectors;
  if (out_scancontexts) {
    std::vector<double> tmp(count * cells);
    CSLAM_CUDA(cudaMemcpyAsync(tmp.data(), h->cols + start * cells, count * cells * sizeof(double),
                               cudaMemcpyDeviceToHost, h->stream));
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    for (int64_t b = 0; b < count; ++b)       // stored column by column; the reference is [rings][sectors]
      for (int j = 0; j < h->sectors; ++j)
        for (int r = 0; r < h->rings; ++r)
          out_scancontexts[b * cells + static_cast<int64_t>(r) * h->sectors + j] = tmp[b * cells + static_cast<int64_t>(j) * h->rings + r];
  }
  if (out_ringkeys) {
    std::vector<double> tmp(count * h->rings);
    CSLAM_CUDA(cudaMemcpy2DAsync(tmp.data(), count * sizeof(double), h->rk + start, h->cap * sizeof(double),
                                 count * sizeof(double), h->rings, cudaMemcpyDeviceToHost, h->stream));
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    for (int64_t b = 0; b < count; ++b)
      for (int r = 0; r < h->rings; ++r) out_ringkeys[b * h->rings + r] = tmp[static_cast<int64_t>(r) * count + b];
  }
  return CSLAM_OK;
}

int cslam_sc_search_host(cslam_sc_t* h, const void* queries, int dtype, int nq, int32_t* out_row,
                         double* out_similarity, int32_t* out_yaw_shift, int32_t* out_candidates,
                         double* out_candidate_dist) {
  CSLAM_REQUIRE(h != nullptr, "cslam_sc_search_host: NULL handle");
  CSLAM_REQUIRE(dtype == CSLAM_DTYPE_F32 || dtype == CSLAM_DTYPE_F64, "cslam_sc_search_host: bad dtype %d", dtype);
  CSLAM_REQUIRE(nq >= 0 && nq <= 65535 && (nq == 0 || (queries && out_row && out_similarity)),
                "cslam_sc_search_host: bad arguments");
  if (nq == 0) return CSLAM_OK;
  CSLAM_REQUIRE(h->n > 0, "cslam_sc_search_host: empty pool (the Python layer answers that case, :55-56)");
  DeviceGuard g(h->device);
  const int R = h->rings, S = h->sectors, C = h->ncand;
  const int64_t cells = static_cast<int64_t>(R) * S;
  CSLAM_TRY(h->qcols.reserve(nq * cells));
  CSLAM_TRY(h->qnorm.reserve(static_cast<size_t>(nq) * S));
  CSLAM_TRY(h->qrk.reserve(static_cast<size_t>(nq) * R));
  // query ring keys are ring-major like the pool's: [rings][nq]
  CSLAM_TRY(sc_prepare(h, queries, dtype, nq, h->qcols.p, h->qnorm.p, h->qrk.p, nq, 0));
  // query tiles of up to 8 (one warp each); with fewer queries the warps split the rows instead
  const int qslots = nq >= 8 ? 8 : nq >= 4 ? 4 : nq >= 2 ? 2 : 1;
  const int rparts = kKnnWarps / qslots;
  const int qtiles = (nq + qslots - 1) / qslots;
  int dev_sms = 148;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, h->device);
  // ranges: ~2 waves of CTAs (4 resident per SM), at least four 256-entry steps each (every
  // range adds `rparts` candidate lists per query to the merge)
  int ranges = (2 * dev_sms * 4 + qtiles - 1) / qtiles;
  const int64_t max_ranges = (h->n + 4 * kKnnRows - 1) / (4 * kKnnRows);
  if (ranges > max_ranges) ranges = static_cast<int>(max_ranges);
  if (ranges < 1) ranges = 1;
  int64_t rows_per_range = (h->n + ranges - 1) / ranges;
  rows_per_range = (rows_per_range + kKnnRows - 1) / kKnnRows * kKnnRows;
  ranges = static_cast<int>((h->n + rows_per_range - 1) / rows_per_range);
  const int parts = ranges * rparts;
  CSLAM_TRY(h->part_d.reserve(static_cast<size_t>(nq) * parts * C));
  CSLAM_TRY(h->part_i.reserve(static_cast<size_t>(nq) * parts * C));
  CSLAM_TRY(h->cand_d.reserve(static_cast<size_t>(nq) * C));
  CSLAM_TRY(h->cand_i.reserve(static_cast<size_t>(nq) * C));
  CSLAM_TRY(h->dist.reserve(static_cast<size_t>(nq) * C));
  CSLAM_TRY(h->yaw.reserve(static_cast<size_t>(nq) * C));
  CSLAM_TRY(h->sim.reserve(nq));
  CSLAM_TRY(h->best_row.reserve(nq));
  CSLAM_TRY(h->best_yaw.reserve(nq));
  cudaEventRecord(h->ev[0], h->stream);
  {
    dim3 grid(ranges, qtiles);
    const size_t smem = static_cast<size_t>(R) * kKnnRows * sizeof(double);
    if (R == 20)
      k_sc_knn<20><<<grid, kKnnThreads, smem, h->stream>>>(h->rk, h->cap, h->n, R, h->qrk.p, nq, C, rows_per_range,
                                                          qslots, h->part_d.p, h->part_i.p);
    else
      k_sc_knn<0><<<grid, kKnnThreads, smem, h->stream>>>(h->rk, h->cap, h->n, R, h->qrk.p, nq, C, rows_per_range,
                                                         qslots, h->part_d.p, h->part_i.p);
    CSLAM_LAUNCH_CHECK();
    k_sc_knn_merge<<<nq, kKnnThreads, 0, h->stream>>>(h->part_d.p, h->part_i.p, parts, C, h->cand_d.p, h->cand_i.p);
    CSLAM_LAUNCH_CHECK();
  }
  cudaEventRecord(h->ev[1], h->stream);
  {
    dim3 grid(C, nq);
    k_sc_distance<<<grid, kDistThreads, dist_smem(R, S), h->stream>>>(h->cols, h->norm, h->qcols.p, h->qnorm.p,
                                                                     h->cand_i.p, C, R, S, h->dist.p, h->yaw.p);
    CSLAM_LAUNCH_CHECK();
    k_sc_pick<<<(nq + 127) / 128, 128, 0, h->stream>>>(h->dist.p, h->yaw.p, h->cand_i.p, nq, C, h->best_row.p,
                                                      h->sim.p, h->best_yaw.p);
    CSLAM_LAUNCH_CHECK();
  }
  cudaEventRecord(h->ev[2], h->stream);
  CSLAM_CUDA(cudaMemcpyAsync(out_row, h->best_row.p, nq * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CSLAM_CUDA(cudaMemcpyAsync(out_similarity, h->sim.p, nq * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (out_yaw_shift)
    CSLAM_CUDA(cudaMemcpyAsync(out_yaw_shift, h->best_yaw.p, nq * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  if (out_candidates)
    CSLAM_CUDA(cudaMemcpyAsync(out_candidates, h->cand_i.p, static_cast<size_t>(nq) * C * sizeof(int32_t),
                               cudaMemcpyDeviceToHost, h->stream));
  if (out_candidate_dist)
    CSLAM_CUDA(cudaMemcpyAsync(out_candidate_dist, h->dist.p, static_cast<size_t>(nq) * C * sizeof(double),
                               cudaMemcpyDeviceToHost, h->stream));
  CSLAM_CUDA(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&h->knn_ms, h->ev[0], h->ev[1]);
  cudaEventElapsedTime(&h->dist_ms, h->ev[1], h->ev[2]);
  return CSLAM_OK;
}

int cslam_sc_last_timing(cslam_sc_t* h, float* knn_ms, float* distance_ms) {
  CSLAM_REQUIRE(h != nullptr, "cslam_sc_last_timing: NULL handle");
  if (knn_ms) *knn_ms = h->knn_ms;
  if (distance_ms) *distance_ms = h->dist_ms;
  return CSLAM_OK;
}

}  // extern "C"
