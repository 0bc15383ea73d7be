// Microbenchmark (measurement tool, not product code): how many SM cycles does one tcgen05.mma kind::f16 take
// on B200 when a single thread issues R of them back to back, as a function of
//   cta_group (1 | 2), N (64 | 128 | 256), A operand source (TMEM ".ts" | shared ".ss"),
//   and whether consecutive MMAs accumulate into the same D columns or alternate between two buffers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_microbench scripts/umma_microbench.cu
// Output: one line per configuration: issue cycles per MMA, completion cycles per MMA (first issue -> commit observed).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

template <int CG, bool TS>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_t, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  if constexpr (TS) {
    if constexpr (CG == 1)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                   "r"(a_t), "l"(b_desc), "r"(idesc), "r"(acc)
                   : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                   "r"(a_t), "l"(b_desc), "r"(idesc), "r"(acc)
                   : "memory");
  } else {
    if constexpr (CG == 1)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                   "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
                   : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                   "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
                   : "memory");
  }
}

// R MMAs, 4 per 64-wide K block (descriptor +2 per K step), cycling through NB K blocks of the B image.
template <int CG, bool TS>
__global__ void __launch_bounds__(128, 1) bench_kernel(int N, int R, int alt, int nb, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t *base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int rows = N / CG;   // B rows per CTA per K block (128 B each)
  const int b_bytes = nb * rows * 128;
  uint32_t *w = reinterpret_cast<uint32_t *>(base);
  for (int i = threadIdx.x; i < (b_bytes + 16384) / 4; i += blockDim.x) w[i] = 0x3c003c00u + (i * 2654435761u & 0x00ff00ffu);
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = s_tmem;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
  const uint32_t blo = ((smem_u32(base) >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t alo = (((smem_u32(base) + b_bytes) >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t kbs = (uint32_t)(rows * 128) >> 4;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (threadIdx.x == 0 && crank == 0) {
    t0 = clock64();
    // 16 MMAs per outer iteration (4 K blocks x 4 K steps), compile-time offsets: the issue stream is UTCHMMA-dense
    for (int r0 = 0; r0 < R; r0 += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint64_t bd = ((uint64_t)DESC_HI << 32) | (uint64_t)(blo + (uint32_t)(j >> 2) * kbs + (uint32_t)(j & 3) * 2u);
        const uint64_t ad = ((uint64_t)DESC_HI << 32) | (uint64_t)(alo + (uint32_t)(j & 3) * 2u);
        const uint32_t d = tb + ((alt && (j & 1)) ? 256u : 0u);
        const uint32_t a_t = tb + ((alt && (j & 1)) ? 0u : 256u) + (uint32_t)(j & 7) * 8u;
        mma<CG, TS>(d, a_t, ad, bd, idesc, (r0 + j) > 1 ? 1u : 0u);
      }
    }
    if constexpr (CG == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                       smem_u32(&bar)),
                   "h"((uint16_t)1)
                   : "memory");
    t1 = clock64();
    while (!mbar_try_wait(smem_u32(&bar), 0)) {}
    t2 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x < 32) {
    if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
  }
}

template <int CG, bool TS>
void run(int N, int R, int alt, int nb, int grid, long long *d_out) {
  const int smem = 1024 + nb * (N / CG) * 128 + 16384 + 64;
  cudaFuncSetAttribute(bench_kernel<CG, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  long long h[2] = {0, 0};
  for (int it = 0; it < 3; ++it) {
    cudaMemset(d_out, 0, 16);
    cudaError_t e = cudaLaunchKernelEx(&cfg, bench_kernel<CG, TS>, N, R, alt, nb, d_out);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("cg=%d ts=%d N=%d: CUDA error %s\n", CG, (int)TS, N, cudaGetErrorString(e));
      exit(1);
    }
    cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
  }
  printf("cg=%d A=%s N=%3d R=%3d alt=%d kblocks=%d grid=%3d : issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)\n", CG,
         TS ? "tmem" : "smem", N, R, alt, nb, grid, (double)h[0] / R, (double)h[1] / R, 128 * N / 256);
}

int main() {
  long long *d_out;
  cudaMalloc(&d_out, 16);
  const int Ns[3] = {64, 128, 256};
  for (int grid : {128})
    for (int alt = 0; alt < 2; ++alt)
      for (int ni = 0; ni < 3; ++ni) {
        const int N = Ns[ni];
        run<1, true>(N, 128, alt, 4, grid, d_out);
        run<1, false>(N, 128, alt, 4, grid, d_out);
        run<2, true>(N, 128, alt, 4, grid, d_out);
        run<2, false>(N, 128, alt, 4, grid, d_out);
      }
  // the rollout kernel's case: cg=2, A in TMEM, B image of a whole 256x256 layer (4 K blocks + bias block)
  for (int N : {128, 256}) run<2, true>(N, 32 * (256 / N), 0, 4, 128, d_out);
  for (int N : {32, 64, 96, 128, 160, 192, 224, 256}) run<2, true>(N, 128, 0, 4, 2, d_out);
  return 0;
}
