// PCA projection on the 5th-gen tensor cores: the GEMM of sklearn `pca.transform`
// (reference cslam/vpr/netvlad.py:234) as a split-K tcgen05.mma kind::tf32 kernel.
//
//   part[z][b][d] = sum_{k in slice z} x[b][k] * W[d][k]        b < 128, d tile of 256
//
// Both operands are K-major float32 in global memory exactly as the reference holds them
// (x = VLAD vectors [B, 32768], W = components_ [4096, 32768]); TMA stages 32-float (128 B)
// k-blocks with the 128B swizzle, the tensor core reads them as TF32 (fp32 accumulate in
// TMEM).  W is 537 MB and is read ONCE per batch: the op is HBM-bound (82 us floor at
// 6.55 TB/s) - the SIMT fp32 kernel it replaces was compute-bound at ~1 ms.
// TF32 operand rounding moves a projected, L2-normalised descriptor element by < 5e-5
// (north_star tolerance 1e-3; tests/test_heads_gpu.py).
//
// Grid: (dout / 256 tiles) x ksplit CTAs so that ~all SMs stream a K-slice of W; the
// partial sums go to the split-K workspace and k_pca_finish (heads.cu) reduces them in a
// fixed order, subtracts mean.W^T, whitens and L2-normalises.
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator,
// warps 2-5 epilogue (tcgen05.ld, warp % 4 = TMEM lane quarter).
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace cslam {
using namespace tc;
namespace {

constexpr int PM = 128;   // batch rows per tile (UMMA M); rows >= batch are zero-filled by TMA
constexpr int PN = 256;   // outputs per tile (UMMA N)
constexpr int PK = 32;    // floats per k-block = one 128B swizzle row
constexpr int P_UMMA_K = 8;
constexpr int P_STAGES = 4;
constexpr int PA_BYTES = PM * PK * 4;   // 16 KiB
constexpr int PB_BYTES = PN * PK * 4;   // 32 KiB
constexpr int P_STAGE_BYTES = PA_BYTES + PB_BYTES;
constexpr int P_THREADS = 192;
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + 1024 + 128;

// kind::tf32 instruction descriptor: D = f32 (c_format 1 at [4,6)), A = B = TF32 (format 2 at
// [7,10) and [10,13)), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) |
                                (static_cast<uint32_t>(PN >> 3) << 17) |
                                (static_cast<uint32_t>(PM >> 4) << 24);

__global__ void __launch_bounds__(P_THREADS, 1)
k_pca_gemm_tc(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
              int batch, int dout, int num_kb, int kb_per_split, float* __restrict__ part) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + P_STAGES * P_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (P_STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * P_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * P_STAGES + 1);
  auto smem_a = [&](int s) { return smem_base + s * P_STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * P_STAGE_BYTES + PA_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * PN;
  const int kz = blockIdx.y;
  const int kb0 = kz * kb_per_split;
  const int kb1 = min(num_kb, kb0 + kb_per_split);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(tmem_slot), "r"(PN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_expect_tx(full_bar(stage), P_STAGE_BYTES);
        tma_load_2d(smem_a(stage), &tmap_x, full_bar(stage), kb * PK, 0, kEvictLast);
        tma_load_2d(smem_b(stage), &tmap_w, full_bar(stage), kb * PK, n0, kEvictFirst);
        if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < PK / P_UMMA_K; ++k) {
          const uint64_t adesc = make_sw128_desc(smem_a(stage) + k * (P_UMMA_K * 4));
          const uint64_t bdesc = make_sw128_desc(smem_b(stage) + k * (P_UMMA_K * 4));
          umma_tf32(tmem_base, adesc, bdesc, kIdescTf32, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
        if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(done_bar);   // accumulator complete
    }
  } else {
    // epilogue: thread = batch row of its TMEM lane quarter; 32 columns per tcgen05.ld
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    float* dst = part + (static_cast<size_t>(kz) * batch + m) * dout + n0;
    const bool has_work = kb1 > kb0;
#pragma unroll 1
    for (int c = 0; c < PN; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(lane_taddr + static_cast<uint32_t>(c), r);
      tmem_ld_wait();
      if (m < batch) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (n0 + c + j + 3 < dout) {
            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                   __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            if (!has_work) v = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(dst + c + j) = v;
          } else {
            for (int e = 0; e < 4; ++e)
              if (n0 + c + j + e < dout) dst[c + j + e] = has_work ? __uint_as_float(r[j + e]) : 0.f;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(PN)
                 : "memory");
  }
}

int make_f32_rowmajor_tmap(CUtensorMap* out, const void* base, int64_t rows, int64_t cols,
                           int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return CSLAM_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 4};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(PK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(f32) failed (CUresult %d) rows=%lld cols=%lld",
              static_cast<int>(r), static_cast<long long>(rows), static_cast<long long>(cols));
    return CSLAM_ERR_CUDA;
  }
  return CSLAM_OK;
}

}  // namespace

// True when the tensor-core path applies: TMA needs 16-byte aligned row strides and bases.
bool pca_tc_supported(const float* d_x, const float* d_w, int din, int dout) {
  return din % 4 == 0 && din >= PK && dout >= 8 && dout % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_w) & 15) == 0;
}

// Split-K partial products for `batch` <= 128 rows; returns the number of K slices written
// (part is [ksplit][batch][dout]) through *ksplit_out.  max_ksplit bounds the workspace.
int launch_pca_gemm_tc(const float* d_x, int batch, int din, const float* d_w, int dout,
                       int max_ksplit, int num_sms, float* d_part, int* ksplit_out,
                       cudaStream_t stream) {
  CUtensorMap tmx, tmw;
  CSLAM_TRY(make_f32_rowmajor_tmap(&tmx, d_x, batch, din, PM));
  CSLAM_TRY(make_f32_rowmajor_tmap(&tmw, d_w, dout, din, PN));
  const int num_kb = (din + PK - 1) / PK;
  const int ntile = (dout + PN - 1) / PN;
  int ksplit = std::max(1, std::min(std::min(max_ksplit, num_kb), std::max(1, num_sms / ntile)));
  const int kb_per = (num_kb + ksplit - 1) / ksplit;
  ksplit = (num_kb + kb_per - 1) / kb_per;   // no empty slices
  CSLAM_CUDA(cudaFuncSetAttribute(k_pca_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  P_SMEM_BYTES));
  dim3 grid(ntile, ksplit);
  k_pca_gemm_tc<<<grid, P_THREADS, P_SMEM_BYTES, stream>>>(tmx, tmw, batch, dout, num_kb, kb_per,
                                                          d_part);
  CSLAM_LAUNCH_CHECK();
  *ksplit_out = ksplit;
  return CSLAM_OK;
}

}  // namespace cslam
