// What bounds "load two operands from shared memory, do a block of independent DFMAs" loops on one SM?
// Variants of the iLQR line-search main loop, one CTA, W warps; prints cycles per loop iteration (thread 0).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/lds_dfma_microbench scripts/lds_dfma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int LA = 10, LSP = 10, WS = 65, KIN = 64, ROWS = 64;

// mode 0: w LDS.64 (row per lane) + 5 uniform LDS.128 + LA DFMA (1 row)      mode 1: + second row (2 LDS.64, 2 LA DFMA)
// mode 2: like 0 but activations as 10 uniform LDS.64                          mode 3: like 0 without the DFMAs
// mode 4: like 0 without the activation loads (h in registers)
template <int MODE>
__global__ void k(double *out, long long *cyc, int iters) {
  extern __shared__ __align__(16) double sm[];
  double *W = sm, *h = sm + ROWS * WS + 1 + ((ROWS * WS + 1) & 1);
  for (int t = threadIdx.x; t < ROWS * WS + KIN * LSP + 8; t += blockDim.x) sm[t] = 1.0 + 1e-3 * t;
  __syncthreads();
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, kq = wp & 3;
  const double *w0p = W + (size_t)(lane) * WS, *w1p = W + (size_t)(lane + 32) * WS;
  double p0[LA], p1[LA];
#pragma unroll
  for (int q = 0; q < LA; ++q) { p0[q] = 0.0; p1[q] = 0.0; }
  double hreg[LA];
#pragma unroll
  for (int q = 0; q < LA; ++q) hreg[q] = h[q];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int kk = kq; kk < KIN; kk += 4) {
      const double w0 = w0p[kk];
      double w1 = 0.0;
      if (MODE == 1) w1 = w1p[kk];
      double hv[LA];
      if (MODE == 2) {
#pragma unroll
        for (int q = 0; q < LA; ++q) hv[q] = h[(size_t)kk * LSP + q];
      } else if (MODE == 4) {
#pragma unroll
        for (int q = 0; q < LA; ++q) hv[q] = hreg[q];
      } else {
        const double2 *h2 = reinterpret_cast<const double2 *>(h + (size_t)kk * LSP);
#pragma unroll
        for (int q = 0; q < LA / 2; ++q) { const double2 v = h2[q]; hv[2 * q] = v.x; hv[2 * q + 1] = v.y; }
      }
      if (MODE == 3) {
#pragma unroll
        for (int q = 0; q < LA; ++q) p0[q] += hv[q] + w0;     // keeps the loads alive with DADDs only on 1 accumulator each
      } else {
#pragma unroll
        for (int q = 0; q < LA; ++q) { p0[q] = fma(w0, hv[q], p0[q]); if (MODE == 1) p1[q] = fma(w1, hv[q], p1[q]); }
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int q = 0; q < LA; ++q) s += p0[q] + p1[q];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char *what, int warps) {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 1024 * sizeof(double)); cudaMalloc(&cyc, 8);
  const int iters = 200;
  const size_t smem = (ROWS * WS + KIN * LSP + 64) * sizeof(double);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<MODE><<<1, warps * 32, smem>>>(out, cyc, iters);
  k<MODE><<<1, warps * 32, smem>>>(out, cyc, iters);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-58s warps=%2d : %.1f cycles per k-iteration (%d per warp per pass)\n", what, warps, (double)h / (iters * (KIN / 4)), KIN / 4);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("1 row:  LDS.64 w + 5 uniform LDS.128 + 10 DFMA", w);
    run<1>("2 rows: 2 LDS.64 w + 5 uniform LDS.128 + 20 DFMA", w);
    run<2>("1 row:  LDS.64 w + 10 uniform LDS.64 + 10 DFMA", w);
    run<3>("1 row:  loads only (DADD instead of DFMA)", w);
    run<4>("1 row:  LDS.64 w + 10 DFMA, activations in registers", w);
  }
  return 0;
}
