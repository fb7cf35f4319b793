// Descriptor-extraction kernels around the PyTorch backbone:
//   A1  k_resample_h / k_resample_v   uint8 HWC keyframe -> center crop -> Pillow-exact
//                                     antialiased bicubic resize -> ToTensor -> Normalize
//                                     (reference: cslam/vpr/netvlad.py:202-208,223-226,
//                                      cslam/vpr/cosplace.py:73-79)
//   A3  k_vlad                        NetVLADLayer.forward fused: per-location L2 norm, 1x1
//                                     conv soft-assignment, softmax, residual aggregation,
//                                     intra-normalisation, global L2   (netvlad.py:94-130)
//   A4  k_pca_gemm / k_pca_finish     sklearn PCA.transform (+whiten) and row L2 normalise
//                                     (netvlad.py:234-237)
//   A5  k_gem_head                    L2Norm -> GeM -> Flatten -> Linear -> L2Norm
//                                     (cslam/vpr/cosplace_utils/network.py:23-29, layers.py:8-36)
// All fp32 like the reference's torch path; these stages are HBM/latency bound
// (25.7 MFLOP and 0.5 MB per image for the VLAD head), so the work is in fusing each stage
// into one pass over its input, not in tensor-core throughput.
#include <math.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace cslam {
bool vlad_tc_supported(const float* d_x, int channels, int locations, int clusters);
int launch_vlad_tc(const float* d_x, int batch, int locations, const float* d_conv_w,
                   const float* d_centroids, float* d_out, cudaStream_t stream);
bool pca_tc_supported(const float* d_x, const float* d_w, int din, int dout);
int launch_pca_gemm_tc(const float* d_x, int batch, int din, const float* d_w, int dout,
                       int max_ksplit, int num_sms, float* d_part, int* ksplit_out,
                       cudaStream_t stream);
namespace {

// ------------------------------------------------------------------ A1 preprocessing
// Pillow's ImagingResample for 8-bit images (libImaging/Resample.c): separable, horizontal
// pass first, coefficients normalised in double then converted to fixed point with
// PRECISION_BITS = 32 - 8 - 2, accumulators start at 1 << (PRECISION_BITS - 1), results
// clipped to uint8 between the passes.  Reproducing that integer arithmetic makes the
// resized image bit-identical to PIL's (tests/test_heads_gpu.py).
constexpr int PRECISION_BITS = 32 - 8 - 2;

double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

struct ResampleTable {
  int ksize = 0;
  std::vector<int> bounds;  // [out][2] = (xmin, count)
  std::vector<int> coef;    // [out][ksize] fixed point
};

ResampleTable precompute_coeffs(int in_size, double in0, double in1, int out_size) {
  ResampleTable t;
  const double support0 = 2.0;  // bicubic
  const double scale = (in1 - in0) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = support0 * filterscale;
  t.ksize = static_cast<int>(ceil(support)) * 2 + 1;
  t.bounds.assign(static_cast<size_t>(out_size) * 2, 0);
  t.coef.assign(static_cast<size_t>(out_size) * t.ksize, 0);
  std::vector<double> k(t.ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = in0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = 0; x < t.ksize; ++x) {
      const double v = x < xmax ? k[x] : 0.0;
      t.coef[static_cast<size_t>(xx) * t.ksize + x] =
          v < 0 ? static_cast<int>(-0.5 + v * (1 << PRECISION_BITS))
                : static_cast<int>(0.5 + v * (1 << PRECISION_BITS));
    }
    t.bounds[2 * xx] = xmin;
    t.bounds[2 * xx + 1] = xmax;
  }
  return t;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass over the cropped window: tmp[b][y][x][c], y in [0, crop_h), x in [0, out)
__global__ void k_resample_h(const uint8_t* __restrict__ img, int B, int H, int W, int top,
                             int left, int crop_h, int out, int ksize,
                             const int* __restrict__ bounds, const int* __restrict__ coef,
                             uint8_t* __restrict__ tmp) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(B) * crop_h * out;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % out);
  const int y = static_cast<int>((idx / out) % crop_h);
  const int b = static_cast<int>(idx / (static_cast<int64_t>(out) * crop_h));
  const int xmin = bounds[2 * x], cnt = bounds[2 * x + 1];
  const uint8_t* row = img + ((static_cast<size_t>(b) * H + top + y) * W + left + xmin) * 3;
  const int* k = coef + static_cast<size_t>(x) * ksize;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int t = 0; t < cnt; ++t) {
    const int kv = k[t];
    s0 += row[3 * t] * kv;
    s1 += row[3 * t + 1] * kv;
    s2 += row[3 * t + 2] * kv;
  }
  uint8_t* o = tmp + static_cast<size_t>(idx) * 3;
  o[0] = clip8(s0);
  o[1] = clip8(s1);
  o[2] = clip8(s2);
}

// vertical pass + ToTensor (/255) + Normalize: out[b][c][y][x] float32
// (Measured and dropped in round 2: both passes in one kernel per 16-row output tile with the input
//  rows and the uint8 intermediate in shared memory - bit-identical, but 0.150 ms instead of 0.104 ms
//  per 64 images: one byte-wide shared-memory load per tap and channel makes it LDS-bound; a
//  register-blocked version, several outputs per thread from words unpacked once, is what it takes.)
__global__ void k_resample_v(const uint8_t* __restrict__ tmp, int B, int crop_h, int out,
                             int ksize, const int* __restrict__ bounds,
                             const int* __restrict__ coef, float m0, float m1, float m2, float d0,
                             float d1, float d2, float* __restrict__ dst) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(B) * out * out;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % out);
  const int y = static_cast<int>((idx / out) % out);
  const int b = static_cast<int>(idx / (static_cast<int64_t>(out) * out));
  const int ymin = bounds[2 * y], cnt = bounds[2 * y + 1];
  const uint8_t* col = tmp + ((static_cast<size_t>(b) * crop_h + ymin) * out + x) * 3;
  const int* k = coef + static_cast<size_t>(y) * ksize;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int t = 0; t < cnt; ++t) {
    const int kv = k[t];
    const uint8_t* p = col + static_cast<size_t>(t) * out * 3;
    s0 += p[0] * kv;
    s1 += p[1] * kv;
    s2 += p[2] * kv;
  }
  // ToTensor: uint8 -> float32 / 255 ; Normalize: (t - mean) / std, all in float32
  const float f0 = __fdiv_rn(static_cast<float>(clip8(s0)), 255.0f);
  const float f1 = __fdiv_rn(static_cast<float>(clip8(s1)), 255.0f);
  const float f2 = __fdiv_rn(static_cast<float>(clip8(s2)), 255.0f);
  const size_t plane = static_cast<size_t>(out) * out;
  float* o = dst + static_cast<size_t>(b) * 3 * plane + static_cast<size_t>(y) * out + x;
  o[0] = __fdiv_rn(__fsub_rn(f0, m0), d0);
  o[plane] = __fdiv_rn(__fsub_rn(f1, m1), d1);
  o[2 * plane] = __fdiv_rn(__fsub_rn(f2, m2), d2);
}

// ------------------------------------------------------------------ A3 NetVLAD head
// One CTA (512 threads = 16 warps) per image.  C = 512 channels, K = 64 clusters fixed by
// the reference (netvlad.py:174-176); S = H*W locations (196 for 224x224 inputs).
constexpr int VC = 512;
constexpr int VK = 64;
constexpr int V_SMAX = 224;   // locations padded to a multiple of 32
constexpr int V_THREADS = 512;
constexpr int V_CCH = 32;     // channel chunk of the soft-assignment GEMM
constexpr int V_SCH = 8;      // location chunk of the aggregation GEMM

struct VladSmem {
  float logits[VK][V_SMAX];          // soft-assignment (later a[k][s] * inv[s])   56 KB
  float inv[V_SMAX];                 // 1 / max(||x[:, s]||, eps)
  float asum[VK];                    // sum_s a[k][s]
  union {
    struct {
      float xs[V_CCH][V_SMAX];       // x chunk, [c][s]                            28 KB
      float ws[VK][V_CCH + 1];       // conv weight chunk                           8 KB
    } g1;
    float xt[V_SCH][VC + 4];         // x chunk transposed, [s][c]                 16 KB
  } u;
  float red[32];
};

__global__ void __launch_bounds__(V_THREADS, 1)
k_vlad(const float* __restrict__ x, int S, const float* __restrict__ conv_w,
       const float* __restrict__ cent, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char vsm_raw[];
  VladSmem& sm = *reinterpret_cast<VladSmem*>(vsm_raw);
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;  // 0..15: owns clusters 4*warp .. 4*warp+3
  const float* xb = x + static_cast<size_t>(b) * VC * S;

  // (1) per-location inverse norm: F.normalize(x, p=2, dim=1), eps = 1e-12
  for (int s = tid; s < V_SMAX; s += V_THREADS) {
    float acc = 0.f;
    if (s < S)
      for (int c = 0; c < VC; ++c) {
        const float v = xb[static_cast<size_t>(c) * S + s];
        acc = fmaf(v, v, acc);
      }
    sm.inv[s] = s < S ? 1.0f / fmaxf(sqrtf(acc), 1e-12f) : 0.f;
  }
  // (2) logits[k][s] = sum_c w[k][c] * xhat[c][s]: thread tile 4 clusters x 7 locations
  float acc[4][7];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) acc[i][j] = 0.f;
  for (int c0 = 0; c0 < VC; c0 += V_CCH) {
    __syncthreads();
    for (int e = tid; e < V_CCH * V_SMAX; e += V_THREADS) {
      const int c = e / V_SMAX, s = e % V_SMAX;
      sm.u.g1.xs[c][s] = s < S ? xb[static_cast<size_t>(c0 + c) * S + s] : 0.f;
    }
    for (int e = tid; e < VK * V_CCH; e += V_THREADS) {
      const int k = e / V_CCH, c = e % V_CCH;
      sm.u.g1.ws[k][c] = conv_w[static_cast<size_t>(k) * VC + c0 + c];
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < V_CCH; ++c) {
      float wv[4], xv[7];
#pragma unroll
      for (int i = 0; i < 4; ++i) wv[i] = sm.u.g1.ws[4 * warp + i][c];
#pragma unroll
      for (int j = 0; j < 7; ++j) xv[j] = sm.u.g1.xs[c][lane + 32 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 7; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int s = lane + 32 * j;
      sm.logits[4 * warp + i][s] = acc[i][j] * sm.inv[s];
    }
  __syncthreads();
  // (3) softmax over the 64 clusters of every location
  for (int s = tid; s < V_SMAX; s += V_THREADS) {
    if (s < S) {
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
  // asum[k] = sum_s a[k][s] (fixed order: deterministic), then fold 1/||x_s|| into a
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = 4 * warp + i;
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) t += sm.logits[k][lane + 32 * j];
    t = warp_sum(t);
    if (lane == 0) sm.asum[k] = t;
#pragma unroll
    for (int j = 0; j < 7; ++j) sm.logits[k][lane + 32 * j] *= sm.inv[lane + 32 * j];
  }
  __syncthreads();
  // (4) vlad[k][c] = sum_s a[k][s] xhat[c][s]: thread tile 4 clusters x 16 channels
  float v[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) v[i][j] = 0.f;
  for (int s0 = 0; s0 < S; s0 += V_SCH) {
    __syncthreads();
    for (int e = tid; e < VC * V_SCH; e += V_THREADS) {
      const int c = e / V_SCH, s = e % V_SCH;
      sm.u.xt[s][c] = (s0 + s < S) ? xb[static_cast<size_t>(c) * S + s0 + s] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < V_SCH; ++s) {
      float av[4], xv[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = sm.logits[4 * warp + i][s0 + s];
#pragma unroll
      for (int j = 0; j < 16; ++j) xv[j] = sm.u.xt[s][lane + 32 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) v[i][j] = fmaf(av[i], xv[j], v[i][j]);
    }
  }
  // (5) residual to the centroids, intra-normalisation (per cluster), global L2
  float total = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = 4 * warp + i;
    const float as = sm.asum[k];
    float n2 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = lane + 32 * j;
      v[i][j] = fmaf(-cent[static_cast<size_t>(k) * VC + c], as, v[i][j]);
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
  __syncthreads();
  if (lane == 0) sm.red[warp] = total;
  __syncthreads();
  float g = 0.f;
  for (int w2 = 0; w2 < V_THREADS / 32; ++w2) g += sm.red[w2];
  const float rg = 1.0f / fmaxf(sqrtf(g), 1e-12f);
  float* ob = out + static_cast<size_t>(b) * VK * VC;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j)
      ob[static_cast<size_t>(4 * warp + i) * VC + lane + 32 * j] = v[i][j] * rg;
}

// ------------------------------------------------------------------ A4 PCA projection
// (tensor-core GEMM in pca_tc.cu; the SIMT kernel below serves shapes TMA cannot address)
// out_raw[b][d] = sum_j x[b][j] W[d][j]  (split-K partials), B <= 64 rows per launch.
// CTA tile 64 (rows) x 64 (outputs), 256 threads, 4x4 register tile, K step 16.
constexpr int PB = 64, PD = 64, PKS = 16;

__global__ void __launch_bounds__(256)
k_pca_gemm(const float* __restrict__ x, int B, int Din, const float* __restrict__ W, int Dout,
           int ksplit, float* __restrict__ part /*[ksplit][B][Dout]*/) {
  __shared__ float xs[PKS][PB + 4];
  __shared__ float ws[PKS][PD + 4];
  const int d0 = blockIdx.x * PD;
  const int kz = blockIdx.y;
  const int klen = (Din + ksplit - 1) / ksplit;
  const int kbeg = kz * klen, kend = min(Din, kbeg + klen);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // outputs 4*tx.., rows 4*ty..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += PKS) {
    for (int e = threadIdx.x; e < PB * PKS; e += 256) {
      const int r = e / PKS, kk = e % PKS;
      xs[kk][r] = (r < B && k0 + kk < kend) ? x[static_cast<size_t>(r) * Din + k0 + kk] : 0.f;
    }
    for (int e = threadIdx.x; e < PD * PKS; e += 256) {
      const int d = e / PKS, kk = e % PKS;
      ws[kk][d] = (d0 + d < Dout && k0 + kk < kend) ? W[static_cast<size_t>(d0 + d) * Din + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < PKS; ++kk) {
      float xv[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = xs[kk][4 * ty + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv[j] = ws[kk][4 * tx + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 4 * ty + i;
    if (r >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + 4 * tx + j;
      if (d < Dout) part[(static_cast<size_t>(kz) * B + r) * Dout + d] = acc[i][j];
    }
  }
}

// sum the split-K partials (fixed order), subtract mean.W^T, whiten, L2-normalise each row
__global__ void __launch_bounds__(256)
k_pca_finish(const float* __restrict__ part, int B, int Dout, int ksplit,
             const float* __restrict__ bias, const float* __restrict__ scale,
             float* __restrict__ out) {
  __shared__ float sh[8];
  const int b = blockIdx.x;
  float n2 = 0.f;
  for (int d = threadIdx.x; d < Dout; d += blockDim.x) {
    float v = 0.f;
    for (int z = 0; z < ksplit; ++z) v += part[(static_cast<size_t>(z) * B + b) * Dout + d];
    v -= bias[d];
    if (scale) v *= scale[d];
    out[static_cast<size_t>(b) * Dout + d] = v;
    n2 = fmaf(v, v, n2);
  }
  n2 = warp_sum(n2);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = n2;
  __syncthreads();
  float t = 0.f;
  for (int k = 0; k < (blockDim.x >> 5); ++k) t += sh[k];
  // sklearn.preprocessing.normalize: rows with zero norm are left unchanged
  const float nrm = sqrtf(t);
  const float r = nrm > 0.f ? 1.0f / nrm : 1.0f;
  for (int d = threadIdx.x; d < Dout; d += blockDim.x) out[static_cast<size_t>(b) * Dout + d] *= r;
}

// ------------------------------------------------------------------ A5 CosPlace head
// One CTA per image: L2Norm over channels per location, GeM pooling, Linear, L2Norm.
// When the image's feature map fits (ResNet trunks: 512 x 7 x 7 floats = 98 KB) it is staged
// in shared memory with coalesced loads and both reductions run from there with all
// threads; otherwise (VGG16 trunk) the map is read from global memory twice.
__global__ void __launch_bounds__(512)
k_gem_head(const float* __restrict__ x, int C, int S, float p, float eps,
           const float* __restrict__ fc_w, const float* __restrict__ fc_b, int D,
           float* __restrict__ out, int stage_map, float* __restrict__ g_out) {
  extern __shared__ __align__(16) float gsm[];
  float* inv = gsm;            // [S]
  float* g = gsm + S;          // [C]
  float* y = g + C;            // [D]
  float* part = y + D;         // [8][S] partial sums of squares
  float* xs = gsm + (9 * S + C + D + 3) / 4 * 4;   // [C*S] staged feature map (stage_map only), 16-byte aligned
  __shared__ float sh[16];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const float* xb = x + static_cast<size_t>(b) * C * S;
  if (stage_map) {
    // 16-byte loads, four in flight per thread: a one-load-per-iteration loop pays a full memory
    // round trip per iteration (49 of them for a 512 x 7 x 7 map = ~35 us of the old 89 us)
    const int n = C * S;
    if ((reinterpret_cast<uintptr_t>(xb) & 15) == 0 && (n & 3) == 0) {
      const float4* src = reinterpret_cast<const float4*>(xb);
      float4* dst = reinterpret_cast<float4*>(xs);
      const int n4 = n >> 2;
      int e = tid;
      for (; e + 3 * static_cast<int>(blockDim.x) < n4; e += 4 * blockDim.x) {
        const float4 v0 = src[e], v1 = src[e + blockDim.x], v2 = src[e + 2 * blockDim.x], v3 = src[e + 3 * blockDim.x];
        dst[e] = v0;
        dst[e + blockDim.x] = v1;
        dst[e + 2 * blockDim.x] = v2;
        dst[e + 3 * blockDim.x] = v3;
      }
      for (; e < n4; e += blockDim.x) dst[e] = src[e];
    } else {
#pragma unroll 4
      for (int e = tid; e < n; e += blockDim.x) xs[e] = xb[e];
    }
    __syncthreads();
    xb = xs;   // generic pointer into shared memory
  }
  // sum over channels of x^2 per location: thread (location, channel eighth); the partials are
  // combined in a fixed order
  for (int t = tid; t < 8 * S; t += blockDim.x) {
    const int s = t % S, part_id = t / S;
    const int c0 = part_id * ((C + 7) / 8), c1 = min(C, c0 + (C + 7) / 8);
    float acc = 0.f;
    for (int c = c0; c < c1; ++c) {
      const float v = xb[static_cast<size_t>(c) * S + s];
      acc = fmaf(v, v, acc);
    }
    part[part_id * S + s] = acc;
  }
  __syncthreads();
  for (int s = tid; s < S; s += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += part[k * S + s];
    inv[s] = 1.0f / fmaxf(sqrtf(acc), 1e-12f);
  }
  __syncthreads();
  // GeM: mean_s clamp(xhat, eps)^p, then ^(1/p)   (layers.py:8-9)
  const bool cube = p == 3.0f;   // the reference's initial (and usual) exponent
  for (int c = tid; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) {
      const float v = fmaxf(xb[static_cast<size_t>(c) * S + s] * inv[s], eps);
      acc += cube ? v * v * v : powf(v, p);
    }
    const float mean = acc / static_cast<float>(S);
    g[c] = cube ? cbrtf(mean) : powf(mean, 1.0f / p);
    if (g_out) g_out[static_cast<size_t>(b) * C + c] = g[c];
  }
  if (g_out) return;   // batched form: Linear + L2Norm run in k_gem_fc / k_gem_norm (CTA-uniform)
  __syncthreads();
  // Linear: y[d] = fc_w[d, :] . g + fc_b[d]; one warp per output row
  for (int d = warp; d < D; d += nw) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(fc_w[static_cast<size_t>(d) * C + c], g[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) y[d] = acc + fc_b[d];
  }
  __syncthreads();
  float n2 = 0.f;
  for (int d = tid; d < D; d += blockDim.x) n2 = fmaf(y[d], y[d], n2);
  n2 = warp_sum(n2);
  if (lane == 0) sh[warp] = n2;
  __syncthreads();
  float t = 0.f;
  for (int k = 0; k < nw; ++k) t += sh[k];
  const float r = 1.0f / fmaxf(sqrtf(t), 1e-12f);
  for (int d = tid; d < D; d += blockDim.x) out[static_cast<size_t>(b) * D + d] = y[d] * r;
}

// Linear of the CosPlace head for a whole batch (network.py:27): a CTA owns a tile of DT output
// rows x 8 images, stages its rows of the weight matrix (read ONCE per 8 images instead of once
// per image: the per-image kernel above pulled the whole 1 MB matrix through one SM per image,
// 89 us for 64 images) and the 8 pooled vectors in shared memory; one warp per (image, row)
// dot product in the same summation order as the per-image kernel (bit-identical results).
constexpr int GEM_IB = 8;
__global__ void __launch_bounds__(512)
k_gem_fc(const float* __restrict__ g, int B, int C, const float* __restrict__ fc_w,
         const float* __restrict__ fc_b, int D, int DT, float* __restrict__ y) {
  extern __shared__ __align__(16) float fsm[];
  float* sw = fsm;                                  // [DT][C]
  float* sg = fsm + static_cast<size_t>(DT) * C;    // [GEM_IB][C]
  const int d0 = blockIdx.x * DT, b0 = blockIdx.y * GEM_IB;
  const int nd = min(DT, D - d0), nbm = min(GEM_IB, B - b0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  // coalesced staging with several loads in flight per thread (a one-load-per-iteration loop pays
  // a memory round trip per iteration: 45 us for the 64 iterations of a 64 x 512 tile)
  auto stage = [&](float* dst, const float* src, int n) {
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 && (n & 3) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(dst);
      const int n4 = n >> 2, step = blockDim.x;
      int e = tid;
      for (; e + 3 * step < n4; e += 4 * step) {
        const float4 v0 = s4[e], v1 = s4[e + step], v2 = s4[e + 2 * step], v3 = s4[e + 3 * step];
        d4[e] = v0;
        d4[e + step] = v1;
        d4[e + 2 * step] = v2;
        d4[e + 3 * step] = v3;
      }
      for (; e < n4; e += step) d4[e] = s4[e];
    } else {
#pragma unroll 4
      for (int e = tid; e < n; e += blockDim.x) dst[e] = src[e];
    }
  };
  stage(sw, fc_w + static_cast<size_t>(d0) * C, nd * C);
  stage(sg, g + static_cast<size_t>(b0) * C, nbm * C);
  __syncthreads();
  // one warp per output row, all images of the tile at once: the weight row is read once, the
  // eight accumulators are independent chains (per (image, row) the order of the sum is unchanged)
  for (int di = warp; di < nd; di += nw) {
    float acc[GEM_IB];
#pragma unroll
    for (int bi = 0; bi < GEM_IB; ++bi) acc[bi] = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float wv = sw[di * C + c];
#pragma unroll
      for (int bi = 0; bi < GEM_IB; ++bi) acc[bi] = fmaf(wv, sg[bi * C + c], acc[bi]);   // rows >= nbm: stale data, never stored
    }
    const float bias = fc_b[d0 + di];
#pragma unroll
    for (int bi = 0; bi < GEM_IB; ++bi) {
      const float v = warp_sum(acc[bi]);
      if (lane == 0 && bi < nbm) y[static_cast<size_t>(b0 + bi) * D + d0 + di] = v + bias;
    }
  }
}

// final L2Norm of every row, same reduction order as the per-image kernel
__global__ void __launch_bounds__(512) k_gem_norm(int D, float* __restrict__ y) {
  __shared__ float sh[16];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  float* row = y + static_cast<size_t>(b) * D;
  float n2 = 0.f;
  for (int d = tid; d < D; d += blockDim.x) n2 = fmaf(row[d], row[d], n2);
  n2 = warp_sum(n2);
  if (lane == 0) sh[warp] = n2;
  __syncthreads();
  float t = 0.f;
  for (int k = 0; k < nw; ++k) t += sh[k];
  const float r = 1.0f / fmaxf(sqrtf(t), 1e-12f);
  for (int d = tid; d < D; d += blockDim.x) row[d] *= r;
}

}  // namespace
}  // namespace cslam

using namespace cslam;

// grow-only per-device scratch for the pooled vectors of the CosPlace head (the entry point has no
// handle; the reference front end is single-threaded, the mutex only guards the bookkeeping)
namespace cslam {
namespace {
int gem_scratch(size_t floats, float** out) {
  static std::mutex mu;
  static float* buf[64] = {};
  static size_t cap[64] = {};
  int dev = 0;
  CSLAM_CUDA(cudaGetDevice(&dev));
  CSLAM_REQUIRE(dev >= 0 && dev < 64, "gem_head_forward: device ordinal %d not supported", dev);
  std::lock_guard<std::mutex> lock(mu);
  if (floats > cap[dev]) {
    CSLAM_CUDA(cudaDeviceSynchronize());
    if (buf[dev]) cudaFree(buf[dev]);
    buf[dev] = nullptr;
    cap[dev] = 0;
    CSLAM_CUDA(cudaMalloc(reinterpret_cast<void**>(&buf[dev]), floats * sizeof(float)));
    cap[dev] = floats;
  }
  *out = buf[dev];
  return CSLAM_OK;
}
}  // namespace
}  // namespace cslam

struct cslam_preproc {
  int device = 0;
  int in_h = 0, in_w = 0, crop_h = 0, crop_w = 0, top = 0, left = 0, out = 0;
  int ksize_h = 0, ksize_v = 0;
  int *d_bounds_h = nullptr, *d_coef_h = nullptr, *d_bounds_v = nullptr, *d_coef_v = nullptr;
  uint8_t* d_tmp = nullptr;
  int tmp_batch = 0;
  float mean[3] = {0.485f, 0.456f, 0.406f};  // IMAGENET_DEFAULT_MEAN / STD (netvlad.py:24-25)
  float stdv[3] = {0.229f, 0.224f, 0.225f};
};

extern "C" {

int cslam_preproc_create(int in_h, int in_w, int crop, int out_size, int device,
                         cslam_preproc_t** out) {
  CSLAM_REQUIRE(out, "preproc_create: out is NULL");
  *out = nullptr;
  CSLAM_REQUIRE(in_h > 0 && in_w > 0 && crop > 0 && out_size > 0, "preproc_create: bad sizes");
  // torchvision CenterCrop pads when the image is smaller than the crop; keyframes here are
  // 640x480 with crop 376 (config/cslam/example.yaml) - padding is not implemented
  CSLAM_REQUIRE(crop <= in_h && crop <= in_w, "preproc_create: crop %d larger than the %dx%d image", crop, in_w, in_h);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("preproc_create: no CUDA device available (this library has no CPU fallback)");
    return CSLAM_ERR_CUDA;
  }
  CSLAM_REQUIRE(device >= 0 && device < ndev, "preproc_create: device %d out of range", device);
  DeviceGuard g(device);
  cslam_preproc* h = new cslam_preproc();
  h->device = device;
  h->in_h = in_h;
  h->in_w = in_w;
  h->crop_h = h->crop_w = crop;
  // torchvision.transforms.functional.center_crop: int(round((H - crop) / 2.0))
  h->top = static_cast<int>(nearbyint((in_h - crop) / 2.0));
  h->left = static_cast<int>(nearbyint((in_w - crop) / 2.0));
  // Resize(out) on a square crop -> out x out
  h->out = out_size;
  ResampleTable th = precompute_coeffs(crop, 0.0, crop, out_size);
  ResampleTable tv = precompute_coeffs(crop, 0.0, crop, out_size);
  h->ksize_h = th.ksize;
  h->ksize_v = tv.ksize;
  int st = CSLAM_OK;
  if ((st = dev_alloc(&h->d_bounds_h, th.bounds.size())) || (st = dev_alloc(&h->d_coef_h, th.coef.size())) ||
      (st = dev_alloc(&h->d_bounds_v, tv.bounds.size())) || (st = dev_alloc(&h->d_coef_v, tv.coef.size()))) {
    cslam_preproc_destroy(h);
    return st;
  }
  cudaMemcpy(h->d_bounds_h, th.bounds.data(), th.bounds.size() * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_coef_h, th.coef.data(), th.coef.size() * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_bounds_v, tv.bounds.data(), tv.bounds.size() * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_coef_v, tv.coef.data(), tv.coef.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (cudaGetLastError() != cudaSuccess) {
    set_error("preproc_create: table upload failed");
    cslam_preproc_destroy(h);
    return CSLAM_ERR_CUDA;
  }
  *out = h;
  return CSLAM_OK;
}

int cslam_preproc_destroy(cslam_preproc_t* h) {
  if (!h) return CSLAM_OK;
  DeviceGuard g(h->device);
  dev_free(h->d_bounds_h);
  dev_free(h->d_coef_h);
  dev_free(h->d_bounds_v);
  dev_free(h->d_coef_v);
  dev_free(h->d_tmp);
  delete h;
  return CSLAM_OK;
}

int cslam_preproc_run(cslam_preproc_t* h, const uint8_t* d_images, int batch, float* d_out,
                      void* stream) {
  CSLAM_REQUIRE(h && d_images && d_out && batch >= 0, "preproc_run: bad arguments");
  if (batch == 0) return CSLAM_OK;
  DeviceGuard g(h->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (batch > h->tmp_batch) {
    CSLAM_CUDA(cudaStreamSynchronize(s));
    dev_free(h->d_tmp);
    CSLAM_TRY(dev_alloc(&h->d_tmp, static_cast<size_t>(batch) * h->crop_h * h->out * 3));
    h->tmp_batch = batch;
  }
  const int64_t t1 = static_cast<int64_t>(batch) * h->crop_h * h->out;
  k_resample_h<<<static_cast<unsigned int>((t1 + 255) / 256), 256, 0, s>>>(
      d_images, batch, h->in_h, h->in_w, h->top, h->left, h->crop_h, h->out, h->ksize_h,
      h->d_bounds_h, h->d_coef_h, h->d_tmp);
  CSLAM_LAUNCH_CHECK();
  const int64_t t2 = static_cast<int64_t>(batch) * h->out * h->out;
  k_resample_v<<<static_cast<unsigned int>((t2 + 255) / 256), 256, 0, s>>>(
      h->d_tmp, batch, h->crop_h, h->out, h->ksize_v, h->d_bounds_v, h->d_coef_v, h->mean[0],
      h->mean[1], h->mean[2], h->stdv[0], h->stdv[1], h->stdv[2], d_out);
  CSLAM_LAUNCH_CHECK();
  return CSLAM_OK;
}

int cslam_vlad_forward(const float* d_x, int batch, int channels, int locations,
                       const float* d_conv_w, const float* d_centroids, int clusters,
                       float* d_out, void* stream) {
  CSLAM_REQUIRE(d_x && d_conv_w && d_centroids && d_out && batch >= 0, "vlad_forward: NULL argument");
  CSLAM_REQUIRE(channels == VC && clusters == VK,
                "vlad_forward: built for %d channels x %d clusters (NetVLAD/VGG16), got %d x %d", VC,
                VK, channels, clusters);
  CSLAM_REQUIRE(locations >= 1 && locations <= V_SMAX, "vlad_forward: 1 <= locations <= %d (got %d)",
                V_SMAX, locations);
  if (batch == 0) return CSLAM_OK;
  // default: aggregation on the tensor cores (vlad_tc.cu); CSLAM_VLAD_TC=0 or a shape TMA cannot
  // address selects the fused fp32 kernel below
  static const bool tc_off = getenv("CSLAM_VLAD_TC") && atoi(getenv("CSLAM_VLAD_TC")) == 0;
  if (!tc_off && vlad_tc_supported(d_x, channels, locations, clusters))
    return launch_vlad_tc(d_x, batch, locations, d_conv_w, d_centroids, d_out,
                          static_cast<cudaStream_t>(stream));
  CSLAM_CUDA(cudaFuncSetAttribute(k_vlad, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(sizeof(VladSmem))));
  k_vlad<<<batch, V_THREADS, sizeof(VladSmem), static_cast<cudaStream_t>(stream)>>>(
      d_x, locations, d_conv_w, d_centroids, d_out);
  CSLAM_LAUNCH_CHECK();
  return CSLAM_OK;
}

int64_t cslam_pca_workspace_floats(int batch, int dout) {
  const int ksplit = 16;
  return static_cast<int64_t>(ksplit) * batch * dout;
}

int cslam_pca_project_l2(const float* d_x, int batch, int din, const float* d_w,
                         const float* d_bias, const float* d_scale, int dout, float* d_out,
                         float* d_work, void* stream) {
  CSLAM_REQUIRE(d_x && d_w && d_bias && d_out && d_work, "pca_project_l2: NULL argument");
  CSLAM_REQUIRE(batch >= 0 && din > 0 && dout > 0, "pca_project_l2: bad sizes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    CSLAM_CUDA(cudaGetDevice(&dev));
    CSLAM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  static const bool force_simt = getenv("CSLAM_PCA_SIMT") != nullptr;
  for (int b0 = 0; b0 < batch; b0 += PB) {
    const int nb = std::min(PB, batch - b0);
    const float* xb = d_x + static_cast<size_t>(b0) * din;
    int ksplit = 16;
    if (!force_simt && pca_tc_supported(xb, d_w, din, dout)) {
      // tcgen05 kind::tf32 split-K GEMM (pca_tc.cu): W streams from HBM once
      CSLAM_TRY(launch_pca_gemm_tc(xb, nb, din, d_w, dout, 16, num_sms, d_work, &ksplit, s));
    } else {
      dim3 grid((dout + PD - 1) / PD, ksplit);
      k_pca_gemm<<<grid, 256, 0, s>>>(xb, nb, din, d_w, dout, ksplit, d_work);
      CSLAM_LAUNCH_CHECK();
    }
    k_pca_finish<<<nb, 256, 0, s>>>(d_work, nb, dout, ksplit, d_bias, d_scale,
                                    d_out + static_cast<size_t>(b0) * dout);
    CSLAM_LAUNCH_CHECK();
  }
  return CSLAM_OK;
}

int cslam_gem_head_forward(const float* d_x, int batch, int channels, int locations, float p,
                           float eps, const float* d_fc_w, const float* d_fc_b, int dout,
                           float* d_out, void* stream) {
  CSLAM_REQUIRE(d_x && d_fc_w && d_fc_b && d_out, "gem_head_forward: NULL argument");
  CSLAM_REQUIRE(batch >= 0 && channels > 0 && locations > 0 && dout > 0, "gem_head_forward: bad sizes");
  if (batch == 0) return CSLAM_OK;
  size_t smem = (static_cast<size_t>(locations) * 9 + channels + dout + 4) * sizeof(float);
  const size_t map_bytes = static_cast<size_t>(channels) * locations * sizeof(float);
  const int stage_map = smem + map_bytes <= 200 * 1024 ? 1 : 0;
  if (stage_map) smem += map_bytes;
  CSLAM_REQUIRE(smem <= 200 * 1024, "gem_head_forward: shapes need %zu B of shared memory", smem);
  CSLAM_CUDA(cudaFuncSetAttribute(k_gem_head, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // batches: pooled vectors to a scratch buffer, then the Linear for 8 images per weight tile and
  // the row normalisation (CSLAM_GEM_FUSED=1 or a single image: everything in the per-image kernel)
  static const bool per_image = getenv("CSLAM_GEM_FUSED") && atoi(getenv("CSLAM_GEM_FUSED")) != 0;
  const int dt = static_cast<int>(std::min<size_t>(64, (160 * 1024) / (static_cast<size_t>(channels) * sizeof(float)) > GEM_IB
                                                           ? (160 * 1024) / (static_cast<size_t>(channels) * sizeof(float)) - GEM_IB
                                                           : 0));
  if (batch > 1 && !per_image && dt >= 1) {
    float* d_g = nullptr;
    CSLAM_TRY(gem_scratch(static_cast<size_t>(batch) * channels, &d_g));
    k_gem_head<<<batch, 512, smem, s>>>(d_x, channels, locations, p, eps, d_fc_w, d_fc_b, dout, d_out, stage_map, d_g);
    CSLAM_LAUNCH_CHECK();
    const size_t fsm_bytes = static_cast<size_t>(dt + GEM_IB) * channels * sizeof(float);
    CSLAM_CUDA(cudaFuncSetAttribute(k_gem_fc, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fsm_bytes)));
    const dim3 grid((dout + dt - 1) / dt, (batch + GEM_IB - 1) / GEM_IB);
    k_gem_fc<<<grid, 512, fsm_bytes, s>>>(d_g, batch, channels, d_fc_w, d_fc_b, dout, dt, d_out);
    CSLAM_LAUNCH_CHECK();
    k_gem_norm<<<batch, 512, 0, s>>>(dout, d_out);
    CSLAM_LAUNCH_CHECK();
    return CSLAM_OK;
  }
  k_gem_head<<<batch, 512, smem, s>>>(d_x, channels, locations, p, eps, d_fc_w, d_fc_b, dout, d_out, stage_map, nullptr);
  CSLAM_LAUNCH_CHECK();
  return CSLAM_OK;
}

}  // extern "C"
