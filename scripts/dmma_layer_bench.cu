// Microbenchmark (measurement tool, not product code): one "layer" of ls_rollouts_mma in isolation (8 warps, one row tile each; 512 threads resident).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_layer_bench scripts/dmma_layer_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts128(uint32_t a, double x, double y) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory"); }
template <int KS, int NW>
__global__ void k(int reps, long long *cyc, double *out) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  for (int i = tid; i < 16384; i += blockDim.x) sm[i] = 1e-3 * (i % 7);
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t S8 = 160, hA = base, hB = base + 64 * S8, W = hB + 64 * S8;
  uint32_t hin = hA, hout = hB;
  const long long t0 = clock64();
  if (wp < NW)
  for (int r = 0; r < reps; ++r) {
    const uint32_t wa = W + (uint32_t)(wp * KS * 32 + lane) * 8u, hb = hin + (uint32_t)(t4 * 20 + g) * 8u;
    double acc[4][2];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = 0.0;
    double a_cur[4], b_cur[4], a_nxt[4], b_nxt[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { a_cur[q] = lds64(wa + q * 256u); b_cur[q] = lds64(hb + q * 4u * S8); }
#pragma unroll 1
    for (int ks = 0; ks < KS; ks += 4) {
      const bool more = ks + 4 < KS;
      if (more) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { a_nxt[q] = lds64(wa + (ks + 4 + q) * 256u); b_nxt[q] = lds64(hb + (ks + 4 + q) * 4u * S8); }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) dmma(acc[q][0], acc[q][1], a_cur[q], b_cur[q]);
      if (more) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { a_cur[q] = a_nxt[q]; b_cur[q] = b_nxt[q]; }
      }
    }
    const int j = 8 * wp + g;
    double y0 = 0.5 + ((acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]));
    double y1 = 0.5 + ((acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]));
    y0 = y0 > 0.0 ? y0 : 0.0; y1 = y1 > 0.0 ? y1 : 0.0;
    sts128(hout + (uint32_t)(j * 20 + 2 * t4) * 8u, y0 * 1e-3, y1 * 1e-3);
    asm volatile("bar.sync 2, %0;" ::"n"(NW * 32) : "memory");
    const uint32_t t2 = hin; hin = hout; hout = t2;
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  out[tid] = sm[tid];
}
template <int KS, int NW> void run() {
  long long *cyc, h; double *out;
  cudaMalloc(&cyc, 8); cudaMalloc(&out, 8 * 1024);
  cudaFuncSetAttribute(k<KS, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
  for (int w = 0; w < 2; ++w) k<KS, NW><<<1, 512, 16384 * 8>>>(1000, cyc, out);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("KS=%2d main warps=%d (512 threads resident): %.0f cycles per layer  (%s)\n", KS, NW, h / 1000.0, cudaGetErrorString(cudaGetLastError()));
}
int main() { run<4, 8>(); run<16, 8>(); run<4, 4>(); run<16, 4>(); run<16, 1>(); return 0; }
