// Microbenchmark (measurement tool, not product code): mma.sync.m8n8k4.f64 on B200 -- latency of a dependent chain,
// issue interval of one warp with ILP independent accumulators, and the SM-wide rate with W warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_microbench scripts/dmma_microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma(int reps, double *out, long long *cyc) {
  double c0[ILP], c1[ILP];
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int q = 0; q < ILP; ++q) c0[q] = c1[q] = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int q = 0; q < ILP; ++q) dmma(c0[q], c1[q], a, b);
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < ILP; ++q) s += c0[q] + c1[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
__global__ void k_dfma(int reps, double *out, long long *cyc) {
  double c[ILP];
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
#pragma unroll
  for (int q = 0; q < ILP; ++q) c[q] = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int q = 0; q < ILP; ++q) c[q] = fma(a, c[q], b);
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < ILP; ++q) s += c[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int warps, bool fma_) {
  double *out;
  long long *cyc, h;
  cudaMalloc(&out, 1024 * 8);
  cudaMalloc(&cyc, 8);
  const int reps = 2000;
  for (int w = 0; w < 2; ++w) {
    if (fma_) k_dfma<ILP><<<1, warps * 32>>>(reps, out, cyc); else k_dmma<ILP><<<1, warps * 32>>>(reps, out, cyc);
  }
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / reps;
  if (fma_) printf("DFMA  warps=%2d ILP=%d : %.1f cycles per round, %.2f cycles per warp-instruction SM-wide, %.1f FMA/clk/SM\n", warps, ILP, per,
                   per / (ILP * warps), 32.0 * ILP * warps / per);
  else printf("DMMA  warps=%2d ILP=%d : %.1f cycles per round, %.2f cycles per DMMA SM-wide, %.1f FMA/clk/SM\n", warps, ILP, per,
              per / (ILP * warps), 256.0 * ILP * warps / per);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<1>(1, false); run<2>(1, false); run<4>(1, false); run<8>(1, false);
  run<1>(4, false); run<4>(4, false); run<8>(4, false); run<8>(8, false); run<8>(16, false);
  run<1>(1, true); run<8>(1, true); run<8>(4, true); run<8>(8, true); run<8>(16, true);
  return 0;
}
