// Internal declarations shared by the NNS translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cslam {

constexpr int kCoarseBM = 128;  // queries per coarse tile
constexpr int kCoarseBN = 256;  // pool rows per coarse tile

// Arguments common to the two candidate-generation kernels (tensor-core coarse
// pass and exact fp64 scan): append (score, row) with score >= tau[q] to query
// q's candidate list.
struct CoarseParams {
  int num_kb;           // dim_pad / 64                     (tensor-core pass only)
  int n_rows;           // valid pool rows
  int num_tiles;        // tiles of kCoarseBN rows to visit
  int tile_stride;      // tile t starts at row t * tile_stride * kCoarseBN
  int nq;               // valid queries in this query tile (<= kCoarseBM)
  int q_row0;           // first row of the query tile in the fp16 query buffer
  const float* tau;     // [nq] thresholds or nullptr (= -inf: keep everything)
  unsigned int* cnt;    // [nq] candidate counters
  uint2* cand;          // [nq, cand_cap] (score bits, row)
  int cand_cap;
};

// 128-byte CUtensorMap blob, kept opaque outside nns_coarse_tc.cu.
struct alignas(64) TensorMapBlob {
  unsigned char bytes[128];
};

int make_fp16_rowmajor_tmap(void* out_map, const void* base, int64_t rows, int cols_pad,
                            int box_rows);
int launch_coarse_tc(const void* tmap_q, const void* tmap_p, const CoarseParams& prm, int num_sms,
                     cudaStream_t stream);

}  // namespace cslam
