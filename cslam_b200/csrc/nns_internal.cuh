// Internal declarations shared by the NNS translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cslam {

constexpr int kCoarseBM = 128;  // queries per coarse tile
constexpr int kCoarseBN = 256;  // pool rows per coarse tile
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
  int nq;               // valid queries in this query tile (<= kCoarseBM)
  int q_row0;           // first row of the query tile in the fp16 query buffer
  const float* tau;     // [nq] thresholds or nullptr (= -inf: keep everything)
  unsigned int* cnt;    // [nq, nsub + 1] per-sub-segment counts; [nsub] = overflow counter
  uint2* cand;          // [nq, nsub * kSegCap + kOvfCap] (score bits, row)
  int nsub;             // sub-segments per query (>= 2 * grid size of the generating kernel)
  uint32_t* smax;       // mode 1: [nq, smax_stride] chunk maxima as ordered keys
  int smax_stride;
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
int launch_coarse_tc(const void* tmap_q, const void* tmap_p, const CoarseParams& prm, int grid,
                     cudaStream_t stream);

}  // namespace cslam
