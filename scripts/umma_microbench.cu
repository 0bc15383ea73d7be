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

// ---------------------------------------------------------------------------------------------------
// Epilogue primitives: cycles per tcgen05.ld 32x32b.x32 (+wait::ld) and per tcgen05.st x16 (+wait::st) as a
// function of the number of warps per SM sub-partition doing it concurrently, with the tensor pipe idle or
// busy with N=128 MMAs accumulating into other TMEM columns (the rollout kernel's situation).
__global__ void __launch_bounds__(32 * 17, 1) epi_bench(int wps, int mma_on, int R, int dcol, int acol, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t *base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint32_t *w = reinterpret_cast<uint32_t *>(base);
  for (int i = threadIdx.x; i < (4 * 128 * 128) / 4; i += blockDim.x) w[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = s_tmem;
  if (warp == 16) {
    if (mma_on && lane == 0) {   // keeps the tensor pipe busy: D = columns [256,384), A = columns [384, 448)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t blo = ((smem_u32(base) >> 4) & 0x3FFFu) | (1u << 16);
      const long long m0 = clock64();
      for (int r0 = 0; r0 < 320; r0 += 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint64_t bd = ((uint64_t)DESC_HI << 32) | (uint64_t)(blo + (uint32_t)(j >> 2) * 1024u + (uint32_t)(j & 3) * 2u);
          mma<1, true>(tb + (uint32_t)dcol, tb + (uint32_t)acol + (uint32_t)(j & 7) * 8u, 0ull, bd, idesc, (r0 + j) > 0 ? 1u : 0u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      while (!mbar_try_wait(smem_u32(&bar), 0)) {}
      if (blockIdx.x == 0) out[4] = clock64() - m0;
    }
  } else if (warp < 4 * wps) {
    const uint32_t la = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) & 1) * 64);   // columns [0,128)
    uint32_t r[32];
    long long t0 = clock64();
    for (int it = 0; it < R; ++it) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
          "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(la + (uint32_t)((it & 1) * 32))
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int q = 0; q < 32; ++q) acc ^= r[q];
    for (int it = 0; it < R; ++it) {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
          "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(la + (uint32_t)((it & 1) * 16)),
          "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
          : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    long long t2 = clock64();
    // two loads in flight before one wait
    for (int it = 0; it < R; ++it) {
      uint32_t q2[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
          "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(la)
          : "memory");
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
          "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(q2[0]), "=r"(q2[1]), "=r"(q2[2]), "=r"(q2[3]), "=r"(q2[4]), "=r"(q2[5]), "=r"(q2[6]), "=r"(q2[7]), "=r"(q2[8]),
            "=r"(q2[9]), "=r"(q2[10]), "=r"(q2[11]), "=r"(q2[12]), "=r"(q2[13]), "=r"(q2[14]), "=r"(q2[15]), "=r"(q2[16]),
            "=r"(q2[17]), "=r"(q2[18]), "=r"(q2[19]), "=r"(q2[20]), "=r"(q2[21]), "=r"(q2[22]), "=r"(q2[23]), "=r"(q2[24]),
            "=r"(q2[25]), "=r"(q2[26]), "=r"(q2[27]), "=r"(q2[28]), "=r"(q2[29]), "=r"(q2[30]), "=r"(q2[31])
          : "r"(la + 32u)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 32; ++q) acc ^= r[q] ^ q2[q];
    }
    long long t3 = clock64();
    if (warp == 0 && lane == 0 && blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t1;
      out[2] = t3 - t2;
      out[3] = (long long)acc;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

void run_epi(int wps, int mma_on, int dcol, int acol, long long *d_out) {
  const int smem = 1024 + 4 * 128 * 128 + 64, R = 200;
  cudaFuncSetAttribute(epi_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[5] = {0, 0, 0, 0, 0};
  for (int it = 0; it < 3; ++it) {
    cudaMemset(d_out, 0, 40);
    epi_bench<<<128, 32 * 17, smem>>>(wps, mma_on, R, dcol, acol, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("epi_bench: CUDA error %s\n", cudaGetErrorString(e));
      exit(1);
    }
    cudaMemcpy(h, d_out, 40, cudaMemcpyDeviceToHost);
  }
  printf("epilogue(ld/st cols [0,128)) warps/SMSP=%d mma=%d D@%d A@%d : ld.x32+wait %.0f cyc, st.x16+wait %.0f cyc, 2 x ld.x32 + 64 LOP + wait %.0f cyc; concurrent N=128 MMA %.1f cyc/MMA\n", wps, mma_on, dcol, acol,
         (double)h[0] / R, (double)h[1] / R, (double)h[2] / R, (double)h[4] / 320);
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
  cudaMalloc(&d_out, 64);
  for (int wps : {1, 2}) run_epi(wps, 0, 256, 384, d_out);
  for (int wps : {0, 2, 4})
    for (int cfg = 0; cfg < 4; ++cfg) {
      const int dcols[4] = {256, 128, 128, 384}, acols[4] = {384, 384, 0, 0};   // A@0 overlaps the ld/st columns (garbage data is fine)
      run_epi(wps, 1, dcols[cfg], acols[cfg], d_out);
    }
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
