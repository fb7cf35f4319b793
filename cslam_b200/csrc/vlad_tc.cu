// NetVLAD layer (reference cslam/vpr/netvlad.py:94-130) with the VLAD aggregation on the 5th-gen
// tensor cores.  Three kernels per batch:
//
//   k_vlad_assign   (fp32 SIMT)  per (image, tile of 128 locations): per-location L2 norm (:105-106),
//                   1x1 conv logits (:109), softmax over the 64 clusters (:110)
//                      a'[b][k][s] = softmax_k(logits)[s] / ||x[:, s]||      (the normalisation of x
//                                                                               folded into a)
//                      asum_part[b][tile][k] = sum_{s in tile} a[k][s]
//   k_vlad_agg_tc   (tcgen05 kind::tf32) per (image, 128 channels):
//                      D[c][k] = sum_s x[b][c][s] * a'[b][k][s]              (:119-124, the contraction)
//                   M = 128 channels, N = 64 clusters, K = locations; x and a' are K-major fp32 in
//                   global memory exactly as they lie (rows of `locations` floats), staged by TMA in
//                   32-float k-blocks with the 128B swizzle (the K tail past `locations` is zero-filled
//                   by TMA), accumulators in TMEM; epilogue: tcgen05.ld, residual to the centroids
//                   v[k][c] = D[c][k] - cent[k][c] * asum[k], written cluster-major (:127 flatten).
//   k_vlad_finish   (fp32 SIMT)  per image: intra-normalisation over channels (:126), global L2 (:128).
// Measured at batch 64 (profiles/r1_vlad_tc_ncu_full.txt): 92 + 13 + 10 us against 225 us for the fused
// fp32 kernel.
//
// Why the soft-assignment GEMM is NOT on the tensor cores: with trained NetVLAD weights the
// 1x1-conv logits are large (conv.weight = 2*alpha*centroids, alpha = 100 in the published
// checkpoints) and the softmax is sharp; a TF32 operand rounding of 2^-11 on logits of that size
// moves the assignment weights by per cent, far outside the 1e-3 descriptor tolerance.  The
// aggregation is linear in its operands, so TF32 costs ~1e-6 absolute on the normalised output
// (tests/test_heads_gpu.py holds it to the same golden vectors as the fp32 kernel).
//
// Shapes TMA cannot address (locations not a multiple of 4, unaligned bases) use the fused fp32
// kernel k_vlad in heads.cu.
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace cslam {
using namespace tc;
namespace {

constexpr int VC = 512;          // channels
constexpr int VK = 64;           // clusters
constexpr int A_STILE = 128;     // locations per assign CTA
constexpr int A_THREADS = 512;   // 16 warps: warp owns 4 clusters, lane owns 4 locations
constexpr int A_CCH = 32;        // channel chunk

struct AssignSmem {
  float logits[VK][A_STILE];     // 32 KB
  float xs[A_CCH][A_STILE];      // 16 KB
  float ws[VK][A_CCH + 1];       //  8 KB
  float nrm2[A_STILE];
  float inv[A_STILE];
};

__global__ void __launch_bounds__(A_THREADS, 1)
k_vlad_assign(const float* __restrict__ x, int S, const float* __restrict__ conv_w,
              float* __restrict__ a_out /*[B][VK][S]*/, float* __restrict__ asum_part /*[B][tiles][VK]*/) {
  extern __shared__ __align__(16) unsigned char asm_raw[];
  AssignSmem& sm = *reinterpret_cast<AssignSmem*>(asm_raw);
  const int tile = blockIdx.x, b = blockIdx.y, tiles = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s0 = tile * A_STILE;
  const int cnt = min(A_STILE, S - s0);
  const float* xb = x + static_cast<size_t>(b) * VC * S;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float n2 = 0.f;    // threads < A_STILE: sum_c x[c][s]^2 of their location, channel order
  // The next channel chunk is fetched into registers while the current one is being multiplied
  // (the chunks were loaded and consumed strictly in turn before: 16 exposed memory round trips
  // per CTA).  Same arithmetic in the same order.
  constexpr int XU = A_CCH * A_STILE / A_THREADS, WU = VK * A_CCH / A_THREADS;
  static_assert(A_CCH * A_STILE % A_THREADS == 0 && VK * A_CCH % A_THREADS == 0, "chunk sizes");
  float rx[XU], rw[WU];
  auto fetch = [&](int c0) {
#pragma unroll
    for (int u = 0; u < XU; ++u) {
      const int e = tid + u * A_THREADS, c = e / A_STILE, s = e % A_STILE;
      rx[u] = s < cnt ? xb[static_cast<size_t>(c0 + c) * S + s0 + s] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < WU; ++u) {
      const int e = tid + u * A_THREADS, k = e / A_CCH, c = e % A_CCH;
      rw[u] = conv_w[static_cast<size_t>(k) * VC + c0 + c];
    }
  };
  fetch(0);
  for (int c0 = 0; c0 < VC; c0 += A_CCH) {
    __syncthreads();
#pragma unroll
    for (int u = 0; u < XU; ++u) {
      const int e = tid + u * A_THREADS;
      sm.xs[e / A_STILE][e % A_STILE] = rx[u];
    }
#pragma unroll
    for (int u = 0; u < WU; ++u) {
      const int e = tid + u * A_THREADS;
      sm.ws[e / A_CCH][e % A_CCH] = rw[u];
    }
    __syncthreads();
    if (c0 + A_CCH < VC) fetch(c0 + A_CCH);
    if (tid < A_STILE) {
#pragma unroll 8
      for (int c = 0; c < A_CCH; ++c) n2 = fmaf(sm.xs[c][tid], sm.xs[c][tid], n2);
    }
#pragma unroll 4
    for (int c = 0; c < A_CCH; ++c) {
      float wv[4], xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) wv[i] = sm.ws[4 * warp + i][c];
#pragma unroll
      for (int j = 0; j < 4; ++j) xv[j] = sm.xs[c][lane + 32 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
    }
  }
  if (tid < A_STILE) sm.inv[tid] = tid < cnt ? 1.0f / fmaxf(sqrtf(n2), 1e-12f) : 0.f;   // F.normalize eps
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) sm.logits[4 * warp + i][lane + 32 * j] = acc[i][j] * sm.inv[lane + 32 * j];
  __syncthreads();
  // softmax over the clusters of every location
  if (tid < A_STILE) {
    const int s = tid;
    if (s < cnt) {
      float mx = -INFINITY;
      for (int k = 0; k < VK; ++k) mx = fmaxf(mx, sm.logits[k][s]);
      float sum = 0.f;
      for (int k = 0; k < VK; ++k) {
        const float e = expf(sm.logits[k][s] - mx);
        sm.logits[k][s] = e;
        sum += e;
      }
      const float r = 1.0f / sum;
      for (int k = 0; k < VK; ++k) sm.logits[k][s] *= r;
    } else {
      for (int k = 0; k < VK; ++k) sm.logits[k][s] = 0.f;
    }
  }
  __syncthreads();
  // per-tile cluster mass (fixed order), then a' = a / ||x_s|| to global, location-contiguous rows
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = 4 * warp + i;
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) t += sm.logits[k][lane + 32 * j];
    t = warp_sum(t);
    if (lane == 0) asum_part[(static_cast<size_t>(b) * tiles + tile) * VK + k] = t;
    float* dst = a_out + (static_cast<size_t>(b) * VK + k) * S + s0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int s = lane + 32 * j;
      if (s < cnt) dst[s] = sm.logits[k][s] * sm.inv[s];
    }
  }
}

// ---- aggregation GEMM on tcgen05 -------------------------------------------------------------
constexpr int GM = 128;   // channels per tile (UMMA M)
constexpr int GN = VK;    // clusters (UMMA N)
constexpr int GK = 32;    // floats per k-block = one 128B swizzle row
constexpr int G_UMMA_K = 8;
constexpr int G_STAGES = 4;
constexpr int GA_BYTES = GM * GK * 4;   // 16 KiB
constexpr int GB_BYTES = GN * GK * 4;   //  8 KiB
constexpr int G_STAGE_BYTES = GA_BYTES + GB_BYTES;
constexpr int G_THREADS = 192;
constexpr int G_SMEM_BYTES = G_STAGES * G_STAGE_BYTES + 1024 + 128;

// kind::tf32 instruction descriptor (see pca_tc.cu): D f32, A = B = TF32, K-major, N >> 3, M >> 4
constexpr uint32_t kIdescAgg = (1u << 4) | (2u << 7) | (2u << 10) |
                               (static_cast<uint32_t>(GN >> 3) << 17) |
                               (static_cast<uint32_t>(GM >> 4) << 24);

__global__ void __launch_bounds__(G_THREADS, 1)
k_vlad_agg_tc(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_a,
              int num_kb, int tiles, const float* __restrict__ cent, const float* __restrict__ asum_part,
              float* __restrict__ out /*[B][VK][VC] un-normalised residual sums*/) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + G_STAGES * G_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (G_STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * G_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * G_STAGES + 1);
  auto smem_a = [&](int s) { return smem_base + s * G_STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * G_STAGE_BYTES + GA_BYTES; };
  __shared__ float s_asum[VK];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, b = blockIdx.y;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(tmem_slot), "r"(GN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + VK) {   // cluster mass of the image: tiles in order
    const int k = threadIdx.x - 64;
    float t = 0.f;
    for (int i = 0; i < tiles; ++i) t += asum_part[(static_cast<size_t>(b) * tiles + i) * VK + k];
    s_asum[k] = t;
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
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_expect_tx(full_bar(stage), G_STAGE_BYTES);
        tma_load_2d(smem_a(stage), &tmap_x, full_bar(stage), kb * GK, b * VC + mt * GM, kEvictFirst);
        tma_load_2d(smem_b(stage), &tmap_a, full_bar(stage), kb * GK, b * VK, kEvictLast);
        if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < GK / G_UMMA_K; ++k) {
          const uint64_t adesc = make_sw128_desc(smem_a(stage) + k * (G_UMMA_K * 4));
          const uint64_t bdesc = make_sw128_desc(smem_b(stage) + k * (G_UMMA_K * 4));
          umma_tf32(tmem_base, adesc, bdesc, kIdescAgg, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
        if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(done_bar);
    }
  } else {
    // epilogue: thread = channel of its TMEM lane quarter, 64 cluster columns
    const int quarter = warp & 3;
    const int c = mt * GM + quarter * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    float* dst = out + static_cast<size_t>(b) * VK * VC + c;
#pragma unroll 1
    for (int k0 = 0; k0 < GN; k0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(lane_taddr + static_cast<uint32_t>(k0), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int k = k0 + j;
        dst[static_cast<size_t>(k) * VC] =
            fmaf(-cent[static_cast<size_t>(k) * VC + c], s_asum[k], __uint_as_float(r[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(GN)
                 : "memory");
  }
}

// One CTA (512 threads) per image: v[k][:] /= max(||v[k][:]||, eps), then the whole vector / its norm.
__global__ void __launch_bounds__(512)
k_vlad_finish(float* __restrict__ out) {
  __shared__ float red[16];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ob = out + static_cast<size_t>(b) * VK * VC;
  float v[4][16];
  float total = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = 4 * warp + i;
    float n2 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[i][j] = ob[static_cast<size_t>(k) * VC + lane + 32 * j];
      n2 = fmaf(v[i][j], v[i][j], n2);
    }
    n2 = warp_sum(n2);
    const float r = 1.0f / fmaxf(sqrtf(n2), 1e-12f);
    float m2 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[i][j] *= r;
      m2 = fmaf(v[i][j], v[i][j], m2);
    }
    total += warp_sum(m2);
  }
  if (lane == 0) red[warp] = total;
  __syncthreads();
  float g = 0.f;
  for (int w = 0; w < 16; ++w) g += red[w];
  const float rg = 1.0f / fmaxf(sqrtf(g), 1e-12f);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) ob[static_cast<size_t>(4 * warp + i) * VC + lane + 32 * j] = v[i][j] * rg;
}

int make_f32_tmap(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return CSLAM_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 4};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(GK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(vlad) failed (CUresult %d) rows=%lld cols=%lld",
              static_cast<int>(r), static_cast<long long>(rows), static_cast<long long>(cols));
    return CSLAM_ERR_CUDA;
  }
  return CSLAM_OK;
}

}  // namespace

bool vlad_tc_supported(const float* d_x, int channels, int locations, int clusters) {
  return channels == VC && clusters == VK && locations % 4 == 0 && locations >= GK &&
         (reinterpret_cast<uintptr_t>(d_x) & 15) == 0;
}

// Scratch of the tensor-core path (a' and the per-tile cluster masses): one grow-only buffer per
// device, reused by every call.  Calls are stream-ordered; a call on another stream than the
// previous one first waits for that one's last kernel (an event), so the buffer is never shared
// by two batches in flight.  (cudaMallocAsync per call measured 1.3 ms of allocator time.)
struct VladScratch {
  float* p = nullptr;
  size_t floats = 0;
  cudaEvent_t done = nullptr;
  cudaStream_t last = nullptr;
  bool used = false;
};
static VladScratch g_vlad_scratch[64];

int launch_vlad_tc(const float* d_x, int batch, int locations, const float* d_conv_w,
                   const float* d_centroids, float* d_out, cudaStream_t stream) {
  const int S = locations;
  const int tiles = (S + A_STILE - 1) / A_STILE;
  int dev = 0;
  CSLAM_CUDA(cudaGetDevice(&dev));
  CSLAM_REQUIRE(dev >= 0 && dev < 64, "vlad_forward: device ordinal %d out of range", dev);
  VladScratch& sc = g_vlad_scratch[dev];
  const size_t a_floats = static_cast<size_t>(batch) * VK * S;
  const size_t a_floats_pad = (a_floats + 63) / 64 * 64;
  const size_t total = a_floats_pad + static_cast<size_t>(batch) * tiles * VK;
  if (!sc.done) CSLAM_CUDA(cudaEventCreateWithFlags(&sc.done, cudaEventDisableTiming));
  if (sc.used && (sc.last != stream || total > sc.floats)) CSLAM_CUDA(cudaStreamWaitEvent(stream, sc.done, 0));
  if (total > sc.floats) {
    if (sc.used) CSLAM_CUDA(cudaEventSynchronize(sc.done));
    dev_free(sc.p);
    sc.floats = 0;
    CSLAM_TRY(dev_alloc(&sc.p, total));
    sc.floats = total;
  }
  float* d_a = sc.p;
  float* d_asum = sc.p + a_floats_pad;
  CUtensorMap tmx, tma;
  CSLAM_TRY(make_f32_tmap(&tmx, d_x, static_cast<int64_t>(batch) * VC, S, GM));
  CSLAM_TRY(make_f32_tmap(&tma, d_a, static_cast<int64_t>(batch) * VK, S, GN));
  static bool attrs_set[64] = {};   // function attributes are per device
  bool& attrs = attrs_set[dev];
  if (!attrs) {
    CSLAM_CUDA(cudaFuncSetAttribute(k_vlad_assign, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sizeof(AssignSmem))));
    CSLAM_CUDA(cudaFuncSetAttribute(k_vlad_agg_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES));
    attrs = true;
  }
  k_vlad_assign<<<dim3(tiles, batch), A_THREADS, sizeof(AssignSmem), stream>>>(d_x, S, d_conv_w, d_a, d_asum);
  CSLAM_LAUNCH_CHECK();
  const int num_kb = (S + GK - 1) / GK;
  k_vlad_agg_tc<<<dim3(VC / GM, batch), G_THREADS, G_SMEM_BYTES, stream>>>(tmx, tma, num_kb, tiles, d_centroids,
                                                                          d_asum, d_out);
  CSLAM_LAUNCH_CHECK();
  k_vlad_finish<<<batch, 512, 0, stream>>>(d_out);
  CSLAM_LAUNCH_CHECK();
  CSLAM_CUDA(cudaEventRecord(sc.done, stream));
  sc.last = stream;
  sc.used = true;
  return CSLAM_OK;
}

}  // namespace cslam
