// FP64 / FP32 FMA latency and per-SM throughput on the target GPU (one CTA, clock64 around an unrolled FMA stream).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/fp64_microbench scripts/fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int ILP>
__global__ void fma_kernel(T *out, long long *cycles, int iters, T a, T b) {
  T acc[ILP];
#pragma unroll
  for (int q = 0; q < ILP; ++q) acc[q] = (T)(threadIdx.x + q);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < ILP; ++q) acc[q] = acc[q] * a + b;
  }
  const long long t1 = clock64();
  __syncthreads();
  T s = 0;
#pragma unroll
  for (int q = 0; q < ILP; ++q) s += acc[q];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

template <typename T, int ILP>
void run(const char *name, int threads) {
  T *out; long long *cyc, h;
  cudaMalloc(&out, 1024 * sizeof(T));
  cudaMalloc(&cyc, sizeof(long long));
  const int iters = 4096;
  fma_kernel<T, ILP><<<1, threads>>>(out, cyc, iters, (T)1.0000001, (T)1e-9);
  fma_kernel<T, ILP><<<1, threads>>>(out, cyc, iters, (T)1.0000001, (T)1e-9);
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double per_warp_instr = (double)h / ((double)iters * ILP);
  const double lanes_per_clk = (double)threads * iters * ILP / (double)h;
  printf("%s threads=%4d ILP=%2d : %.2f cycles per warp-instruction (per thread stream), %.2f FMA lanes/clk/SM\n", name,
         threads, ILP, per_warp_instr, lanes_per_clk);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("# %s, %d SMs\n", p.name, p.multiProcessorCount);
  run<double, 1>("f64", 32); run<double, 8>("f64", 32); run<double, 8>("f64", 128); run<double, 8>("f64", 640); run<double, 8>("f64", 1024);
  run<float, 1>("f32", 32); run<float, 8>("f32", 32); run<float, 8>("f32", 128); run<float, 8>("f32", 1024);
  return 0;
}
