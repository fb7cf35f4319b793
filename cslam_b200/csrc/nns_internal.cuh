// Internal declarations shared by the NNS translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cslam {

constexpr int kCoarseBM = 128;  // queries per coarse tile
constexpr int kCoarseBN = 256;  // pool rows per coarse tile
constexpr int kMaxCluster = 4;  // query tiles that share one pool sweep (thread-block cluster)
constexpr int kGroupMax = kCoarseBM * kMaxCluster;  // queries per search group
// Candidate lists: per query, `nsub` private sub-segments of kSegCap slots followed by a
// shared overflow region of kOvfCap slots.  Sub-segment (2 * cta + half) is owned by the
// one thread of CTA `cta` that scans column half `half` of every tile for that query, so
// appends need no atomics; only the overflow region is claimed with atomicAdd.
constexpr int kSegCap = 128;
constexpr int kOvfCap = 4096;
constexpr int kChunk = 32;  // columns per epilogue chunk = granularity of the sample maxima

// Arguments common to the two candidate-generation kernels (tensor-core coarse
// pass and exact fp64 scan).
//   mode 0: append (score, row) with score >= tau[q] to query q's candidate list
//   mode 1: sample pass - store the maximum score of every 32-row chunk as an ordered key
struct CoarseParams {
  int mode;
  int num_kb;           // dim_pad / 64                     (tensor-core pass only)
  int n_rows;           // valid pool rows
  int num_tiles;        // tiles of kCoarseBN rows to visit
  int tile_stride;      // tile t starts at row t * tile_stride * kCoarseBN
  int nq;               // valid queries in this group (<= kCoarseBM * cluster)
  int q_row0;           // first row of the group in the fp16 query buffer
  int cluster;          // CTAs per cluster = query tiles per group (1, 2 or 4)
  int l2_prefetch;      // pool tiles pulled into L2 ahead of the shared-memory ring (0 = off)
  int a_resident;       // 1: query tile resident in smem, only pool tiles stream (cluster, dim_pad <= 512)
  const float* tau;     // [nq] thresholds or nullptr (= -inf: keep everything)
  unsigned int* cnt;    // [nq, nsub + 1] per-sub-segment counts; [nsub] = overflow counter
  uint2* cand;          // [nq, nsub * kSegCap + kOvfCap] (score bits, row)
  int nsub;             // sub-segments per query (>= 2 * grid size of the generating kernel)
  uint32_t* smax;       // mode 1: [nq, smax_stride] chunk maxima as ordered keys
  int smax_stride;
  long long* dbg;       // optional [8] cycle counters written by CTA 0 (CSLAM_NNS_DEBUG)
};

__host__ __device__ inline size_t cand_slots(int nsub) {
  return static_cast<size_t>(nsub) * kSegCap + kOvfCap;
}

// 128-byte CUtensorMap blob, kept opaque outside nns_coarse_tc.cu.
struct alignas(64) TensorMapBlob {
  unsigned char bytes[128];
};

int make_fp16_rowmajor_tmap(void* out_map, const void* base, int64_t rows, int cols_pad,
                            int box_rows);
// tmap_p must have box rows kCoarseBN / prm.cluster; grid = CTAs (a multiple of prm.cluster)
int launch_coarse_tc(const void* tmap_q, const void* tmap_p, const CoarseParams& prm, int grid,
                     cudaStream_t stream);
int launch_coarse_pair(const void* tmap_q, const void* tmap_p, const CoarseParams& prm, int grid,
                       cudaStream_t stream);
// co-resident clusters of the given size (cudaOccupancyMaxActiveClusters)
int coarse_tc_max_clusters(int cluster, int* out);

}  // namespace cslam
