// Micro-benchmark: latency / issue cost of the instructions the small dense solves of the
// eigen-solver are made of, for ONE warp on an SM (the situation of the Rayleigh-Ritz step).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* out, long long* cyc, double seed, int n) {
  double a = seed + threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9;
  double a2 = a + 1, a3 = a + 2, a4 = a + 3, a5 = a + 4, a6 = a + 5, a7 = a + 6, a8 = a + 7;
  float f = (float)a;
  __shared__ double sm[256];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (MODE == 0) { a = fma(a, b, c); }                                   // dependent DFMA chain
    if (MODE == 1) { a = fma(a, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c);
                     a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c); a8 = fma(a8, b, c); }  // 8 independent
    if (MODE == 2) { a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31); }   // dependent 64-bit shuffle
    if (MODE == 3) { a = rsqrt(a) + 1.0; }                                  // fp64 rsqrt chain
    if (MODE == 4) { f = rsqrtf(f) + 1.0f; }                                // fp32 MUFU chain
    if (MODE == 5) { f = fmaf(f, 1.0000001f, 1e-9f); }                      // dependent FFMA chain
    if (MODE == 6) { a = sqrt(a) + 1.0; }
    if (MODE == 7) { a = 1.0 / a + 1.0; }
    if (MODE == 8) { int ex; double m = frexp(a, &ex); a = ldexp(m, 3) + 1.0; }
    if (MODE == 9) { a = (double)(float)a + 1e-3; }                         // cvt round trip
    if (MODE == 10) { const int src = (threadIdx.x + 1) & 31;              // 8 independent 64-bit shuffles
                      a = __shfl_sync(0xffffffffu, a, src); a2 = __shfl_sync(0xffffffffu, a2, src);
                      a3 = __shfl_sync(0xffffffffu, a3, src); a4 = __shfl_sync(0xffffffffu, a4, src);
                      a5 = __shfl_sync(0xffffffffu, a5, src); a6 = __shfl_sync(0xffffffffu, a6, src);
                      a7 = __shfl_sync(0xffffffffu, a7, src); a8 = __shfl_sync(0xffffffffu, a8, src); }
    if (MODE == 11) { sm[threadIdx.x] = a; __syncwarp(); a = sm[(threadIdx.x + 1) & 31] + 1e-9; __syncwarp(); }  // smem round trip
    if (MODE == 12) { a = fma(a, b, c); a2 = a2 * b; a3 = a3 + c; a4 = fma(a4, b, c); }   // 4 independent fp64
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + a2 + a3 + a4 + a5 + a6 + a7 + a8 + f;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
  const int n = 4096;
  const char* names[] = {"dfma dependent", "dfma 8 independent (per 8)", "shfl64 dependent", "rsqrt(double)+add",
                         "rsqrtf+add", "ffma dependent", "sqrt(double)+add", "1/x double + add", "frexp+ldexp+add", "cvt f64->f32->f64 + add",
                         "shfl64 8 independent (per 8)", "smem store+syncwarp+load+syncwarp", "4 independent fp64 (per 4)"};
  for (int threads : {32}) {
    for (int mode = 0; mode < 13; ++mode) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (mode) {
#define CASE(M) case M: k<M><<<1, threads>>>(out, cyc, 1.5, n); break;
          CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
        }
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("threads %3d  %-28s %8.1f cycles/iter\n", threads, names[mode], (double)h / n);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
