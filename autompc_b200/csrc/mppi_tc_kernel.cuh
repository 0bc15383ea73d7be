// MPPI rollout + softmax update on the 5th-generation tensor cores (tcgen05 / UMMA, sm_100a).
//
// One launch = one MPPI.run (autompc/control/mppi.py:154-168), same contract as mppi_fp32.cu.
//
// Mapping.  A CTA owns 128 samples; sample t <-> thread t <-> TMEM lane t, so the state, the
// running cost and the noise of a sample never leave its thread.  Per horizon step the MLP is a
// chain of n_layers dependent GEMMs  D[128 x N_l] = A[128 x K_l] . W_l^T :
//   * W_l (torch.nn.Linear layout (out,in) == K-major B operand) is converted to bf16 once on the
//     host, laid out as the UMMA canonical K-major SWIZZLE_128B shared-memory image, and stays
//     RESIDENT in shared memory for the whole solve (no per-step weight traffic);
//   * A_l (the activations) lives in TENSOR MEMORY as packed bf16 and is the TMEM A operand of
//     tcgen05.mma (".ts" form) -- activations never touch shared memory;
//   * D_l is the fp32 accumulator in TMEM; the sample's thread reads its row with tcgen05.ld,
//     applies bias + activation, packs to bf16 and writes the next layer's A row with tcgen05.st.
// With CG=2 the kernel runs as CTA pairs (cta_group::2, UMMA M=256): each CTA keeps HALF of the
// output neurons of every layer in its shared memory, which is what lets the 3x256 network
// (283 KB of bf16 weights) stay resident; the leader CTA's single MMA thread issues for both.
// Warps 0-3 ("control", thread t <-> sample t): Philox noise, clipping, clipped-noise write-back,
// action and control cost -- run one horizon step AHEAD of the GEMM chain through a double-buffered
// shared array handed over with named barriers, so none of it is on the critical path.
// Warps 4-7 ("owners", thread 128+t <-> sample t, normalised state in registers): integration, next input.
// Warps 8-11 ("helpers", thread 256+t <-> sample t): stage cost of the state.  Both groups share the
// layer epilogues: warp w serves TMEM lane quarter w%4 and one 32/64-column half of every N-half.
// Warp 12: MMA issue + TMEM allocation.  (Role order = issue priority: highest warp id first.)
// Pipelining inside the dependent GEMM chain.  Every GEMM is issued at full width (N up to 256: 16
// tcgen05.mma of K=16 per 256-wide layer), accumulating alternately into two 256-column TMEM
// buffers.  The epilogue of GEMM n reads D_n 64 columns at a time and writes the packed bf16
// activations IN PLACE over the accumulator columns it has just consumed; each 64-column group is
// one K-group of GEMM n+1, released to the issuer through its own mbarrier (bar_a[g]), so GEMM n+1
// starts after the FIRST group is packed and only a quarter of each epilogue is exposed.
// bar_d (tcgen05.commit, multicast to the CTA pair) = "accumulator of the GEMM complete".
// Work that does not depend on the state (next step's noise, clipping, control cost) runs in the
// shadow of the layer-1 MMAs.
//
// DZ build (template flag; ReLU networks with >= 2 hidden layers): the state update is linear, z_{i+1} = z_i + dz_i, so the
// input layer of step i+1 is issued as  W0 [z_i | u_{i+1} | 1]  (16-bit operands the owners stored one layer earlier, in
// tensor-memory columns the last hidden GEMM does not read)  +  W0x dz_i  (kind::tf32 MMAs whose A operand is the output
// layer's fp32 accumulator, read in place): no epilogue hop between the output layer and the next input layer, and the
// owners integrate z off the critical path.  The layer loop of the epilogue warps is rolled in that build
// (mppi_tc_hidden_layer.inc is its body in both builds).
//
// The clipped noise (mppi.py:139) of every (step, control, sample) is written to an L2-resident
// scratch and re-read once the softmax weights are known (mppi.py:115-117); per-CTA partial records
// (min, sum w, sum w*eps) are merged by the last CTA to finish, exactly like the fp32 kernel.
#pragma once
#include <cuda_bf16.h>

#include <type_traits>

#include "ampc_common.cuh"

// Kernel template + device helpers.  Included by mppi_tc.cu (host side: plans, weight images, launch) and by the
// mppi_tc_inst_*.cu translation units, each of which instantiates a quarter of the (cta_group, NXP, activation)
// matrix so that the 21 instantiations compile in parallel.
namespace ampc_tc {


constexpr int TM = 128;              // samples per CTA
constexpr int NEPI = 256;            // 8 epilogue warps: 4 owner + 4 helper
constexpr int NCTL = 128;            // 4 control warps (noise / clipping / control cost, one step ahead)
constexpr int NTHR = NEPI + NCTL + 32;   // + 1 MMA warp
constexpr int MMA_WARP = (NEPI + NCTL) / 32;
constexpr int BAR_FULL = 1, BAR_EMPTY = 3;   // named barriers 1,2 / 3,4: hand-over of the two control buffers
constexpr int BAR_X = 5;             // owners -> helpers: the shared copy of the state is up to date
constexpr int TMEM_COLS = 512;
constexpr int TMEM_BUF = 256;        // two accumulator / activation buffers: columns [0,256) and [256,512)
constexpr int MAXL = AMPC_MAX_LAYERS;
constexpr int TRACE_EV = 120;           // (120: the headline shape in dz mode + the timeline ring just fit in 227 KB)
// upper word of the K-major SWIZZLE_128B descriptor: SBO = 1024 B (bits 32-45), version 1 (bit 46), layout 2 (bits 61-63)
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr int DEFER_J = 4;           // K-steps of (half 1, K-pair 0) issued in the first phase; the other 8 - DEFER_J follow commit0
constexpr int YCOL = 64;             // accumulator columns of the output-layer GEMM inside its buffer (clear of the next input block)
constexpr int MAXG = 2;              // N-halves of a GEMM = K-pairs of the next one

struct TcArgs {
  int n_layers;
  int kpad[MAXL], npad[MAXL];        // padded K (16s; hidden: 64s) and N (hidden: 64s; output: 32s) per layer
  uint32_t w_off[MAXL];              // byte offset of layer l's B image inside one CTA's weight image
  uint32_t w_bytes;                  // bytes of one CTA's weight image
  int b_off[MAXL];                   // float offset of the layer's (padded) bias
  int bias_floats;
  uint32_t idesc[MAXL];              // UMMA instruction descriptors (N = chunk width)
  int cw[MAXL], nch[MAXL];           // = hwid, nh (kept for the weight-image row permutation)
  int nh[MAXL], hwid[MAXL];          // N-halves of the layer's GEMM and their width (npad / nh)
  int nkp[MAXL], awid[MAXL];         // K-pairs of the layer's GEMM and their width in K elements
  const uint8_t *wimg;               // CG images back to back
  const float *bias;
  float *epsc;                       // (H*nu, Kc) clipped noise scratch
  int Kc;                            // grid * 128
  int nxp;                           // padded state width (kernel template): input K columns [0,nxp) = state
  int ones[MAXL];                    // layer l's epilogue also writes the constant-one K-step of layer l+1 (bias fold)
  int defer_j;                       // K-steps of (half 1, K-pair 0) issued before the wait for K-pair 1 (2 | 4 | 6; 8 = no deferral)
  // "dz" mode (see the kernel header): layer 0 of step i+1 = 16-bit GEMM on [z_i | u_{i+1} | 1] issued EARLY
  //   + kind::tf32 GEMM whose A operand is the output layer's fp32 accumulator (the increment of z) read in place
  int dz;                            // != 0: on
  uint32_t wx_off;                   // byte offset of the tf32 image of the input layer's state columns (rows x 128 B)
  uint32_t idesc_x;                  // kind::tf32 instruction descriptor (N = hwid[0])
  int nkx;                           // tf32 K-steps (8 states each) = ceil(nxp / 8)
  int ecol;                          // dz mode: TMEM column (inside the buffer) of the early 16-bit input block
  int l0_packed;                     // input layer's image: N-half 1 sits in bytes [64,128) of N-half 0's rows (kpad[0] <= 32)
  unsigned long long *trace;         // debug timeline (AMPC_TC_TRACE=1), else null: [warp][event] = clock<<8 | tag
};

// ------------------------------------------------------------------ PTX wrappers ---
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// arrive on the barrier at the same shared offset in CTA `cta` of the cluster (works for the own CTA too)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {   // non-suspending probe
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {
      printf("ampc mppi_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(TMEM_COLS) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(TMEM_COLS) : "memory");
}
// D[tmem] (+)= A[tmem, packed bf16] . B[smem desc]^T ; executed by ONE thread (the elected issuer).
// b_desc_lo = low word of the K-major SWIZZLE_128B descriptor (start address >> 4 | LBO); the high word is constant.
template <int CG>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_desc_lo, uint32_t idesc,
                                        uint32_t accumulate) {
  const uint64_t b_desc = ((uint64_t)DESC_HI << 32) | (uint64_t)b_desc_lo;
  if constexpr (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same with kind::tf32: A = 8 fp32 TMEM columns per K-step (an fp32 accumulator read in place), B = fp32 K-major image
template <int CG>
__device__ __forceinline__ void umma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_desc_lo, uint32_t idesc,
                                             uint32_t accumulate) {
  const uint64_t b_desc = ((uint64_t)DESC_HI << 32) | (uint64_t)b_desc_lo;
  if constexpr (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t e;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
  return e != 0;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// {lo, hi} -> packed 16-bit pair (element with the even K index in the low half).  F16 = false: bf16 (8-bit
// mantissa, fp32 range); F16 = true: IEEE half (11-bit mantissa -- the operand precision of kind::tf32 -- at the
// full kind::f16 rate; finite range 65504, so the conversions saturate instead of producing inf).
template <bool F16>
__device__ __forceinline__ uint32_t pack_16(float lo, float hi) {
  uint32_t d;
  if constexpr (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <bool F16>
__device__ __forceinline__ uint32_t pack_16_relu(float lo, float hi) {
  uint32_t d;
  if constexpr (F16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// d^T M d for this thread's sample; v laid out [i][TM]
__device__ __forceinline__ float quad_full(const float *M, const float *v, const float *off, int n, bool diag, int t) {
  float c = 0.f;
  if (diag) {
#pragma unroll 8
    for (int i = 0; i < n; ++i) {
      const float di = v[i * TM + t] - (off ? off[i] : 0.f);
      c = fmaf(M[i * n + i] * di, di, c);
    }
  } else {
    for (int i = 0; i < n; ++i) {
      const float di = v[i * TM + t] - (off ? off[i] : 0.f);
      float row = 0.f;
      for (int j = 0; j < n; ++j) row = fmaf(M[i * n + j], v[j * TM + t] - (off ? off[j] : 0.f), row);
      c = fmaf(di, row, c);
    }
  }
  return c;
}

// activation + bf16 pack of NV consecutive accumulator columns (the bias is already in the accumulator:
// it enters every hidden GEMM through a constant-one K-step, split in two bf16 terms)
template <int NV, bool RELU, bool F16>
__device__ __forceinline__ void epi_pack(const uint32_t (&r)[NV], int act, uint32_t (&pk)[NV / 2]) {
  if constexpr (RELU) {
#pragma unroll
    for (int q = 0; q < NV / 2; ++q) pk[q] = pack_16_relu<F16>(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
  } else {
#pragma unroll
    for (int q = 0; q < NV / 2; ++q)
      pk[q] = pack_16<F16>(ampc_act<float>(act, __uint_as_float(r[2 * q])), ampc_act<float>(act, __uint_as_float(r[2 * q + 1])));
  }
}

// Issues the KSP K-steps of one K-pair of one N-half, fully unrolled with compile-time column offsets:
// the pair's K elements sit in two sub-halves (one per epilogue warp of a lane quarter), each packed at the
// start of its own KSP*8 accumulator columns.
template <int CG, int KSP, int J0 = 0, int J1 = KSP>
__device__ __forceinline__ void issue_pair(uint32_t dh, uint32_t a_pair, uint32_t desc_pair, uint32_t kb_stride,
                                           uint32_t idesc, bool first_pair) {
#pragma unroll
  for (int j = J0; j < J1; ++j) {
    const uint32_t acol = (uint32_t)((j / (KSP / 2)) * (KSP * 8) + (j % (KSP / 2)) * 8);
    const uint32_t d = desc_pair + (uint32_t)(j >> 2) * kb_stride + (uint32_t)(j & 3) * 2u;
    umma_ts<CG>(dh, a_pair + acol, d, idesc, (j == 0 && first_pair) ? 0u : 1u);
  }
}

// RELU: the activation is compiled in (the reference's default, mlp.py:44-53); the other activations share one
// instantiation with a runtime switch.  Keeping their code out of the ReLU kernel matters: inlined, it put 43 KB
// of cold instructions between the LDTM, the packs and the STTM of every epilogue (instruction-cache misses on
// the critical path).
template <int CG, int NXP, bool RELU, bool F16, bool TRACE, bool DZ>
__global__ void __launch_bounds__(NTHR, 1) mppi_rollout_tc_kernel(const AmpcMppiParams p, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t t_entry = (uint32_t)clock();
  const int nx = p.nx, nu = p.nu, H = p.H, HN = H * nu, L = a.n_layers;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const AmpcConstLayout cl(nx, nu);

  // ---- shared memory carve (weight image first, 1024-byte aligned for the 128B swizzle atoms)
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t *base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t *s_w = base;
  float *s_const = reinterpret_cast<float *>(base + a.w_bytes);
  float *s_act = s_const + cl.total;                   // shifted act_sequence (H*nu)
  float *s_x = s_act + ((HN + 3) & ~3);                // state [nx][128]
  float *s_u = s_x + nx * TM;                          // scaled control, two buffers of [nu][128]
  float2 *s_zc = reinterpret_cast<float2 *>(s_u + 2 * nu * TM);   // input z-score as (scale, bias) per K column [64]
  float4 *s_qc = reinterpret_cast<float4 *>(s_zc + 64);   // stage cost from z: (std, mean - goal, Q_jj, 0) per state [32]
  float2 *s_xc = reinterpret_cast<float2 *>(s_qc + 32);                            // x = z * std + mean per state [32]
  float *s_wgt = reinterpret_cast<float *>(s_xc + 32); // helper cost share, then softmax numerators [128]
  float *s_cc = s_wgt + TM;                            // control warps' cost share [128]
  float *s_red = s_cc + TM;                            // 32
  float *s_misc = s_red + 32;                          // 64 + AMPC_MERGE_CACHE
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_misc + 64 + AMPC_MERGE_CACHE);   // [0..1]=bar_d[h], [2..3]=bar_a[kp]
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 2 * MAXG + 2);   // s_bar[2*MAXG] = bar_w (weight image landed), [2*MAXG+1] = bar_y
  __shared__ int s_last;

  // Weight image (this CTA's half): TMA bulk copies global -> shared, completion counted in bytes on bar_w.  One
  // thread issues them before anything else so that they overlap the rest of the setup; it waits for them just
  // before the setup barrier.  (The staged LDG/STS loop this replaces was most of the 7.8 k-cycle setup.)
  const uint32_t bar_w = smem_u32(&s_bar[2 * MAXG]);
  if (tid == 0) {
    mbar_init(bar_w, 1);
    fence_mbar_init();
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_w), "r"(a.w_bytes) : "memory");
    const uint8_t *src = a.wimg + (size_t)cta_rank * a.w_bytes;
    constexpr uint32_t CHUNK = 32768;
    for (uint32_t off = 0; off < a.w_bytes; off += CHUNK) {
      const uint32_t n = (a.w_bytes - off < CHUNK) ? a.w_bytes - off : CHUNK;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(s_w + off)),
                   "l"(src + off), "r"(n), "r"(bar_w)
                   : "memory");
    }
  }
  for (int i = tid; i < cl.total; i += NTHR) s_const[i] = p.consts[i];
  for (int e = tid; e < HN; e += NTHR) {               // mppi.py:122-123
    const int i = e / nu, j = e - i * nu;
    const int src = (i + 1 < H) ? i + 1 : H - 1;
    s_act[e] = p.act_seq[src * nu + j];
  }
  // mppi.py:129-130: every sample starts from x0.  One load per state per CTA (x0 may live in mapped host memory:
  // ampc_mppi_solve_host reads the observation zero-copy); s_misc is free until the merge at the end.
  for (int j = tid; j < nx; j += NTHR) s_misc[j] = p.x0_inline ? p.x0_val[j] : p.x0[j];
  // K column k of the input layer: k < NXP -> state k (zero beyond nx); NXP <= k < NXP+nu -> control k-NXP
  // (the weight image uses the same permutation).  z = v * scale + bias  (mlp.py:20-24).
  for (int k = tid; k < 64; k += NTHR) {
    const int j = (k < NXP) ? (k < nx ? k : -1) : (k - NXP < nu ? nx + (k - NXP) : -1);
    float2 zc = make_float2(0.f, 0.f);
    if (j >= 0) {
      const float inv = p.consts[cl.xu_inv + j];
      zc = make_float2(inv, -p.consts[cl.xu_mean + j] * inv);
    }
    if (k == NXP + nu || k == NXP + nu + 1) zc = make_float2(0.f, 1.f);   // constant one: carries the layer-0 bias
    s_zc[k] = zc;
  }
  // The owner keeps the NORMALISED state z = (x - mean) / std in registers (it is what the input layer eats):
  //   x' = x + (y + b_out) * dy_std + dy_mean  (mlp.py:26-30, :236)   <=>   z' = z + y * k1 + k2,
  //   k1 = dy_std / std, k2 = (b_out * dy_std + dy_mean) / std, both folded into the output layer's weight image
  //   (rows scaled by k1, k2 on the constant-one K-step), so the accumulator IS the increment of z;
  //   x = z * std + mean is recovered off the critical path.
  for (int j = tid; j < 32; j += NTHR) {
    float2 xc = make_float2(0.f, 0.f);
    float4 qc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < nx) {
      xc = make_float2(1.f / p.consts[cl.xu_inv + j], p.consts[cl.xu_mean + j]);
      qc = make_float4(xc.x, xc.y - p.consts[cl.goal + j], p.consts[cl.Q + j * nx + j], 0.f);
    }
    s_xc[j] = xc;
    s_qc[j] = qc;
  }
  const uint32_t bar_d0 = smem_u32(&s_bar[0]), bar_a0 = smem_u32(&s_bar[MAXG]);
  const uint32_t bar_y = smem_u32(&s_bar[2 * MAXG + 1]);   // dz mode: "output-layer accumulator complete"
  constexpr bool dz = DZ;
  if (tid == 0) {
    mbar_init(bar_y, 1);
    for (int g = 0; g < MAXG; ++g) mbar_init(bar_d0 + 8u * g, 1);
    for (int g = 0; g < MAXG; ++g) mbar_init(bar_a0 + 8u * g, (NEPI / 32) * CG);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) tmem_alloc<CG>(smem_u32(s_tmem));
  if (tid == 0) mbar_wait(bar_w, 0);                   // weight image landed (async proxy -> async proxy: no fence needed)
  fence_proxy_async_smem();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // debug timeline of CTA 0 (AMPC_TC_TRACE=1): up to TRACE_EV events per warp, steps 10 and 11 plus the kernel
  // phases (step -1), kept in shared memory while the kernel runs (a clock read + one STS per event) and copied
  // to global at the end; event = clock << 8 | tag, tag = kind*16 + index
  uint32_t *s_trace = reinterpret_cast<uint32_t *>(s_tmem + 4);
  int trace_n = 0;
  auto trace = [&](int step, int tag) {
    if constexpr (TRACE) {
      if (blockIdx.x == 0 && lane == 0 && (step < 0 || (step >= 10 && step < 12)) && trace_n < TRACE_EV)
        s_trace[warp * TRACE_EV + trace_n++] = ((uint32_t)clock() << 8) | (uint32_t)(tag & 255);
    }
  };
  trace(-1, 0xA1);                                      // setup done (weights resident, TMEM allocated)
  const float *c_goal = s_const + cl.goal, *c_Q = s_const + cl.Q, *c_R = s_const + cl.R, *c_F = s_const + cl.F;
  const float *c_lo = s_const + cl.lo, *c_hi = s_const + cl.hi, *c_scale = s_const + cl.scale;

  const int t = tid & (TM - 1);                        // sample slot of this thread (owner, helper or control)
  const int k_local = blockIdx.x * TM + t;
  const bool valid = k_local < p.K;
  float cost_acc = 0.f;

  if (warp == MMA_WARP) {
    // =========================== MMA issuer (leader CTA of the pair) ===========================
    // ONE elected thread runs this loop (tcgen05.mma / .commit are single-thread instructions).
    // GEMM n = (step, layer).  Its N is issued as nh halves (one commit each: bar_d[h]); its K as nkp
    // pairs, pair kp being exactly what the epilogue of half kp of GEMM n-1 produced (bar_a[kp]).
    // Issue order (h0,kp0) (h1,kp0) | (h0,kp1)+commit (h1,kp1)+commit keeps the tensor pipe busy with the
    // kp0 work of GEMM n while the epilogue of half 1 of GEMM n-1 is still running.
    // All 512 TMEM columns are allocated, so the base is column 0 / lane 0 (checked below): accumulator and
    // operand addresses are literals, the per-layer descriptor words are hoisted out of the horizon loop, and
    // the K-steps of a pair are unrolled with compile-time offsets -> the SASS is a dense run of UTCHMMA fed by
    // uniform-datapath adds (v8 spent 55-90 cycles per MMA on ELECT/VOTEU/R2UR sequences; a N=128 MMA executes in 66).
    if (cta_rank == 0 && elect_one()) {
      if (tmem_base != 0u) {
        printf("ampc mppi_tc: unexpected TMEM base %u\n", tmem_base);
        __trap();
      }
      uint32_t pa = 0;                                  // parity bit kp of bar_a[kp]
      uint32_t n = 0;                                   // GEMM counter: D_n in buffer n&1, A_n in the other one
      const uint32_t w_addr = smem_u32(s_w);
      uint32_t lo_l[MAXL], kbs_l[MAXL], hro_l[MAXL], id_l[MAXL];
      int nh_l[MAXL], nkp_l[MAXL], ksp_l[MAXL], hw_l[MAXL], aw_l[MAXL];
#pragma unroll
      for (int l = 0; l < MAXL; ++l) {
        const int rows = a.npad[l] / CG;                // B rows held by each CTA (per 64-wide K block)
        lo_l[l] = (((w_addr + a.w_off[l]) >> 4) & 0x3FFFu) | (1u << 16);
        kbs_l[l] = (uint32_t)(rows * 128) >> 4;
        hro_l[l] = (uint32_t)((a.hwid[l] / CG) * 128) >> 4;   // B rows of one N-half (descriptor units)
        if (l == 0 && a.l0_packed) hro_l[l] = 64u >> 4;       // ... or the second 64 bytes of the same rows
        id_l[l] = a.idesc[l];
        nh_l[l] = a.nh[l];
        nkp_l[l] = a.nkp[l];
        ksp_l[l] = a.awid[l] >> 4;                      // K-steps per pair
        hw_l[l] = a.hwid[l];
        aw_l[l] = a.awid[l];
      }
      const int nks0 = a.kpad[0] >> 4;
      const int defer_j = a.defer_j;
      // dz mode: descriptor words of the input layer's two N-halves, 16-bit image (E) and tf32 image of the state columns (T)
      constexpr int NKX = (NXP + 7) / 8;
      const uint32_t xlo = (((w_addr + a.wx_off) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t xhro = (uint32_t)((a.hwid[0] / CG) * 128) >> 4;
      const uint32_t e_lo[MAXG] = {lo_l[0], lo_l[0] + hro_l[0]}, x_lo[MAXG] = {xlo, xlo + xhro};
      const uint32_t hw0 = (uint32_t)hw_l[0], idx = a.idesc_x, ecol = (uint32_t)a.ecol;
      const int nh0 = nh_l[0];
      for (int i = 0; i < H; ++i) {
#pragma unroll
        for (int l = 0; l < MAXL; ++l) {
          if (l >= L) break;
          if constexpr (dz) {
            if (l == 0 && i > 0) continue;              // issued at the end of the previous step (below)
          }
          const int nh = nh_l[l], nkp = nkp_l[l];
          const uint32_t idesc = id_l[l], kb_stride = kbs_l[l];
          const uint32_t d_addr = (n & 1u) * TMEM_BUF + (l == L - 1 ? (uint32_t)YCOL : 0u);
          const uint32_t a_addr = ((n + 1u) & 1u) * TMEM_BUF;
          // Full-width hidden GEMMs (two N-halves, two K-pairs of 8 K-steps) are issued as
          //   (h0,kp0) (h1,kp0: first DEFER_J K-steps) | wait kp1 | (h0,kp1)+commit0 (h1,kp0: the rest) (h1,kp1)+commit1
          // so that 12 MMAs (~800 cycles) are still queued behind commit0 -- the MMA pipeline latency, the epilogue of
          // half 0 and the cross-CTA hand-over (~800 cycles together) finish before the pipe drains -- while the 14
          // MMAs of the first phase cover the wait for kp1.  Measured on one box (AMPC_TC_DEFER = 2 / 4 / 6 / 8):
          // 0.2429 / 0.2404 / 0.2436 / 0.2538 ms per solve at C3.
          const bool defer = (defer_j < 8 && l > 0 && nh == 2 && nkp == 2 && ksp_l[l] == 8);
          for (int kp = 0; kp < nkp; ++kp) {
            mbar_wait(bar_a0 + 8u * kp, (pa >> kp) & 1u);
            pa ^= (1u << kp);
            tc_fence_after();
            trace(i, 0x10 + l * 2 + kp);                // bar_a[kp] of layer l observed
            const uint32_t a_pair = a_addr + (uint32_t)(kp * aw_l[l]);
            const uint32_t pair_off = (uint32_t)((kp * ksp_l[l]) >> 2) * kb_stride;
            for (int h = 0; h < nh; ++h) {
              const uint32_t dh = d_addr + (uint32_t)(h * hw_l[l]);
              const uint32_t hb0 = lo_l[l] + hro_l[l] * (uint32_t)h;
              const uint32_t hb = hb0 + pair_off;
              if (l == 0) {                           // input layer: 1..4 K-steps at columns {0, 8, 32, 40}
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  if (ks < nks0)
                    umma_ts<CG>(dh, a_pair + (uint32_t)((ks >> 1) * 32 + (ks & 1) * 8), hb + (uint32_t)(ks * 2), idesc,
                                ks > 0 ? 1u : 0u);
              } else if (defer && h == 1) {
                if (kp == 0) {
                  if (defer_j == 2) issue_pair<CG, 8, 0, 2>(dh, a_pair, hb, kb_stride, idesc, true);
                  else if (defer_j == 4) issue_pair<CG, 8, 0, 4>(dh, a_pair, hb, kb_stride, idesc, true);
                  else issue_pair<CG, 8, 0, 6>(dh, a_pair, hb, kb_stride, idesc, true);
                } else {                                // rest of K-pair 0
                  if (defer_j == 2) issue_pair<CG, 8, 2, 8>(dh, a_addr, hb0, kb_stride, idesc, false);
                  else if (defer_j == 4) issue_pair<CG, 8, 4, 8>(dh, a_addr, hb0, kb_stride, idesc, false);
                  else issue_pair<CG, 8, 6, 8>(dh, a_addr, hb0, kb_stride, idesc, false);
                  issue_pair<CG, 8>(dh, a_pair, hb, kb_stride, idesc, false);
                }
              } else {
                if (ksp_l[l] == 8) issue_pair<CG, 8>(dh, a_pair, hb, kb_stride, idesc, kp == 0);
                else issue_pair<CG, 4>(dh, a_pair, hb, kb_stride, idesc, kp == 0);
              }
              if (l > 0 && kp == 0)                   // constant-one K-step: the layer's bias (extra K block of the image)
                umma_ts<CG>(dh, a_addr + (uint32_t)(aw_l[l] >> 2), hb0 + (uint32_t)(a.kpad[l] >> 6) * kb_stride, idesc, 1u);
              if (kp == nkp - 1) {                     // half h committed (dz mode: the output layer reports on bar_y)
                umma_commit<CG>((dz && l == L - 1) ? bar_y : bar_d0 + 8u * h);
                trace(i, 0x20 + l * 2 + h);
              }
            }
          }
          ++n;
        }
        if constexpr (dz) if (i + 1 < H) {
          // ---- input layer of step i+1 (GEMM n), dz mode.  z_{i+1} = z_i + dz_i, so
          //   W0 [z_{i+1} | u_{i+1} | 1] = W0 [z_i | u_{i+1} | 1]  (16-bit operands; the owners stored them at TMEM columns
          //                                                         [ecol, ecol + 8 nks0) of this buffer -- columns the last
          //                                                         hidden GEMM does not read -- before their first release
          //                                                         of that layer's epilogue: "early" part E)
          //                              + W0x dz_i               (kind::tf32; A = the output layer's fp32 accumulator, read
          //                                                         in place at YCOL of the same buffer, no epilogue hop: T).
          // Order  E(h0) E(h1) | wait bar_y | T(h0) commit0  T(h1) commit1 :  E executes while this thread waits for the
          // output layer's commit (T reads what other MMAs wrote: not a pair the tensor pipe orders by itself).  (With
          // E(h1) behind commit0 instead: 0.2361 vs 0.2343 ms on one box.)
          const uint32_t d_addr = (n & 1u) * TMEM_BUF, a_addr = ((n + 1u) & 1u) * TMEM_BUF;
#pragma unroll
          for (int h = 0; h < MAXG; ++h) {
            if (h < nh0) {
              const uint32_t dh = d_addr + (uint32_t)h * hw0;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (ks < nks0)
                  umma_ts<CG>(dh, a_addr + ecol + (uint32_t)(ks * 8), e_lo[h] + (uint32_t)(ks * 2), id_l[0], ks > 0 ? 1u : 0u);
            }
          }
          trace(i, 0x61);                               // E issued
          if (a.dz != 2) {                              // (dz == 2: measurement only -- rely on the pipe's issue order)
            if (!mbar_test_wait(bar_y, (uint32_t)i & 1u)) mbar_wait(bar_y, (uint32_t)i & 1u);
            tc_fence_after();
          }
          trace(i, 0x60);                               // output-layer accumulator of step i complete (issuer)
#pragma unroll
          for (int h = 0; h < MAXG; ++h) {
            if (h < nh0) {
              const uint32_t dh = d_addr + (uint32_t)h * hw0;
#pragma unroll
              for (int ks = 0; ks < NKX; ++ks)
                umma_ts_tf32<CG>(dh, a_addr + (uint32_t)(YCOL + ks * 8), x_lo[h] + (uint32_t)(ks * 2), idx, 1u);
              umma_commit<CG>(bar_d0 + 8u * h);
              trace(i + 1, 0x20 + h);
            }
          }
          ++n;
        }
      }
    }
    __syncwarp();
  } else if (warp < NCTL / 32) {
    // =========================== control warps (0-3): one horizon step ahead ===========================
    // Lowest warp ids on purpose: the SM sub-partition's arbiter favours the highest eligible warp id, so the
    // long ALU streams of the noise generator only take issue slots the latency-critical warps (epilogue 4-11,
    // MMA issuer 12) leave free.  (With the control warps on ids 8-11 the owners' and helpers' short critical
    // sections ran 3-5x slower whenever Philox was in flight.)
    const uint32_t kg = (uint32_t)(p.k_offset + k_local);
    const int nblk = (nu + 3) >> 2;
    // controls of step i: noise, clip, write-back (mppi.py:134-139), action cost (:143), control cost (:142)
    auto prepare_controls = [&](int i, float *su) {
      for (int blk = 0; blk < nblk; ++blk) {
        float n4[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.eps == nullptr) {
          ampc_normal4(p.seed, p.ctr, kg, (uint32_t)i, (uint32_t)blk, n4);
#pragma unroll
          for (int q = 0; q < 4; ++q) n4[q] *= p.sqrt_sigma;
        } else if (valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = blk * 4 + q;
            if (j < nu) n4[q] = p.eps[((size_t)i * p.K + k_local) * nu + j];
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = blk * 4 + q;
          if (j < nu) {
            const float a0 = s_act[i * nu + j];
            const float av = fminf(c_hi[j], fmaxf(c_lo[j], n4[q] + a0));
            const float e = av - a0;
            a.epsc[(size_t)(i * nu + j) * a.Kc + k_local] = e;
            su[j * TM + t] = av * c_scale[j];
            cost_acc = fmaf(p.lam_over_sigma * av, e, cost_acc);
          }
        }
      }
      cost_acc += quad_full(c_R, su, nullptr, nu, p.r_diag, t);
    };
    for (int i = 0; i < H; ++i) {
      const int b = i & 1;
      if (i >= 2) asm volatile("bar.sync %0, %1;" ::"r"(BAR_EMPTY + b), "n"(NEPI / 2 + NCTL) : "memory");   // owners read step i-2
      prepare_controls(i, s_u + b * nu * TM);
      __threadfence_block();
      asm volatile("bar.arrive %0, %1;" ::"r"(BAR_FULL + b), "n"(NEPI / 2 + NCTL) : "memory");
    }
    s_cc[t] = cost_acc;                                 // action + control costs of the sample
  } else {
    // =========================== epilogue warps: owners (0-3) and helpers (4-7) ===========================
    // Everything between "accumulator complete" and "activations released" is on the critical path of the
    // dependent GEMM chain, so this code is kept short: layers and halves are unrolled at compile time (their
    // shapes come straight from the constant bank), both TMEM loads of a half are issued before the one wait,
    // and no debug code is compiled in unless TRACE.
    const bool owner = warp < 8;                        // warps 4-7: owners, 8-11: helpers
    const int hf = (warp >> 2) - 1;                     // which 32 of every 64 columns this warp serves
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t pd = 0;                                    // parity bit h of bar_d[h]
    uint32_t n = 0;                                     // GEMM counter (see the issuer)
    auto wait_d = [&](int h) {
      mbar_wait(bar_d0 + 8u * h, (pd >> h) & 1u);
      pd ^= (1u << h);
      tc_fence_after();
    };
    auto signal_a = [&](int g) {                        // "my part of activation group g is in TMEM, my D reads are done"
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(bar_a0 + 8u * g, 0u); else mbar_arrive_local(bar_a0 + 8u * g);
      }
    };
    // the sample's normalised state lives in its owner's registers
    float z[NXP];
#pragma unroll
    for (int j = 0; j < NXP; ++j) {
      const float2 zc = s_zc[j];
      z[j] = (owner && j < nx) ? fmaf(s_misc[j], zc.x, zc.y) : 0.f;
    }
    const int kpad0 = a.kpad[0];
    // Input-layer A operand (bf16, K columns [0,NXP) = z, [NXP,NXP+nu) = z-scored controls, then the constant ones):
    // K-step g (16 columns) -> 8 TMEM columns at {0, 8, 32, 40}.  K-steps that hold no state are written as soon as the
    // controls are known (pack_controls, in the shadow of the output-layer MMAs, whose accumulator sits at YCOL and
    // does not overlap them); the control half of the K-step that straddles NXP waits in cpk; store_input adds the
    // state once y has arrived.
    constexpr int NMIX = (NXP % 16) ? (16 - NXP % 16) / 2 : 1;
    uint32_t cpk[NMIX];
    auto zctl = [&](const float *su, int k) -> float {  // z-scored control / constant-one column k >= NXP
      const float2 zc = s_zc[k];
      return fmaf((k - NXP < nu) ? su[(k - NXP) * TM + t] : 0.f, zc.x, zc.y);
    };
    // controls of `step`: acquire = wait until the control warps have them (named barrier, also a rendezvous of the four
    // owner warps: kept out of critical sections); pack = z-score, pack and store the K-steps that hold no state
    auto acquire_controls = [&](int step) {
      asm volatile("bar.sync %0, %1;" ::"r"(BAR_FULL + (step & 1)), "n"(NEPI / 2 + NCTL) : "memory");
    };
    auto pack_acquired_controls = [&](int step, uint32_t buf) {
      const int b = step & 1;
      const float *su = s_u + b * nu * TM;
      if constexpr (NXP % 16 != 0) {
#pragma unroll
        for (int q = 0; q < NMIX; ++q) cpk[q] = pack_16<F16>(zctl(su, NXP + 2 * q), zctl(su, NXP + 2 * q + 1));
      }
#pragma unroll
      for (int g = (NXP + 15) / 16; g < 4; ++g) {
        if (g * 16 < kpad0) {
          uint32_t pk[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) pk[q] = pack_16<F16>(zctl(su, g * 16 + 2 * q), zctl(su, g * 16 + 2 * q + 1));
          tmem_st8(lane_base + buf * TMEM_BUF + (uint32_t)((g >> 1) * 32 + (g & 1) * 8), pk);
        }
      }
      if (step + 2 < H) asm volatile("bar.arrive %0, %1;" ::"r"(BAR_EMPTY + b), "n"(NEPI / 2 + NCTL) : "memory");
    };
    auto pack_controls = [&](int step, uint32_t buf) {
      acquire_controls(step);
      pack_acquired_controls(step, buf);
    };
    // dz mode: the 16-bit operand of the NEXT step's input layer, [z_i | z-scored u_{i+1} | ones], K-step g at TMEM columns
    // ecol + 8 g of buffer `buf` (columns of the last hidden GEMM's A buffer that its MMAs never read: the 32 columns an
    // owner warp freed when it packed 64 accumulator columns into 32, or columns beyond a narrower layer)
    auto early_input = [&](int step, uint32_t buf) {
      const int b = step & 1;
      const float *su = s_u + b * nu * TM;
      asm volatile("bar.sync %0, %1;" ::"r"(BAR_FULL + b), "n"(NEPI / 2 + NCTL) : "memory");   // controls of `step` are ready
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (g * 16 < kpad0) {
          uint32_t pk[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int k = g * 16 + q * 2;               // NXP is even: a pair is all state or all control
            if (k < NXP) pk[q] = pack_16<F16>(z[k < NXP ? k : 0], z[k + 1 < NXP ? k + 1 : 0]);
            else pk[q] = pack_16<F16>(zctl(su, k), zctl(su, k + 1));
          }
          tmem_st8(lane_base + buf * TMEM_BUF + (uint32_t)a.ecol + (uint32_t)(g * 8), pk);
        }
      }
      if (step + 2 < H) asm volatile("bar.arrive %0, %1;" ::"r"(BAR_EMPTY + b), "n"(NEPI / 2 + NCTL) : "memory");
    };
    auto store_input = [&](uint32_t buf) {
#pragma unroll
      for (int g = 0; g * 16 < NXP; ++g) {
        uint32_t pk[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int k = g * 16 + q * 2;                 // NXP is even: a pair is all state or all control
          if (k < NXP) pk[q] = pack_16<F16>(z[k < NXP ? k : 0], z[k + 1 < NXP ? k + 1 : 0]);
          else pk[q] = cpk[(k - NXP) / 2 < NMIX ? (k - NXP) / 2 : 0];
        }
        tmem_st8(lane_base + buf * TMEM_BUF + (uint32_t)((g >> 1) * 32 + (g & 1) * 8), pk);
      }
    };
    auto store_x = [&]() {
#pragma unroll
      for (int j = 0; j < NXP; ++j)
        if (j < nx) {
          const float2 xc = s_xc[j];
          s_x[j * TM + t] = fmaf(z[j], xc.x, xc.y);
        }
    };
    if (owner) {
      pack_controls(0, 1u);
      store_input(1u);                                  // GEMM 0 reads A from buffer 1
    }
    signal_a(0);

    for (int i = 0; i < H; ++i) {
      // ---- hidden layers: D (buffer n&1) -> activation -> 16 bit, written IN PLACE over the consumed accumulator
      //      columns; each N-half is one K-pair of the next GEMM.  The dz build keeps ONE copy of this loop body (no
      //      unrolling over layers: a third of the instruction footprint, and early_input is not replicated per layer).
      if constexpr (!dz) {
#pragma unroll
        for (int l = 0; l < MAXL - 1; ++l) {
          if (l >= L - 1) break;
#include "mppi_tc_hidden_layer.inc"
        }
      } else {
#pragma unroll 1
        for (int l = 0; l < L - 1; ++l) {
#include "mppi_tc_hidden_layer.inc"
        }
      }
      // ---- output layer: integrate in z space (mlp.py:235-236), then the next step's input in place.
      //      The control columns of that input are packed while the output-layer MMAs run.
      if constexpr (dz) {
        // dz mode: the 16-bit part of the next input was built from z_i before the last hidden epilogue (above); the
        // increment the output layer is computing reaches the next input layer through the tensor pipe (kind::tf32 on
        // the accumulator), so integrating z is off the critical path.
        if (owner) {
          mbar_wait(bar_y, (uint32_t)i & 1u);
          tc_fence_after();
          trace(i, 0x30 + (L - 1) * 2);
          uint32_t r[32];
          tmem_ld32(lane_base + (n & 1u) * TMEM_BUF + YCOL, r);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < NXP; ++j) z[j] += __uint_as_float(r[j]);
          trace(i, 0x91);
        }
      } else {
        if (owner && i + 1 < H) pack_controls(i + 1, n & 1u);
        wait_d(0);
        trace(i, 0x30 + (L - 1) * 2);
        if (owner) {
          const uint32_t dbuf = lane_base + (n & 1u) * TMEM_BUF;
          uint32_t r[32];
          tmem_ld32(dbuf + YCOL, r);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < NXP; ++j) z[j] += __uint_as_float(r[j]);   // the image carries dy_std / std and the constants
          if (i + 1 < H) store_input(n & 1u);           // GEMM n+1 reads A from buffer n&1
        }
        if (i + 1 < H) signal_a(0);
        trace(i, 0x50);                                 // next input released
      }
      ++n;
    }
    // x_H for the terminal cost goes into the same shared array the helpers read the last stage cost from: nothing
    // in the GEMM chain orders the two at the LAST step (a one-hidden-layer network with a dense Q lost the race)
    if (owner) {
      asm volatile("bar.sync %0, %1;" ::"n"(BAR_X), "n"(NEPI) : "memory");
      store_x();
    } else {
      s_wgt[t] = cost_acc;                              // helper's share (state costs)
      asm volatile("bar.arrive %0, %1;" ::"n"(BAR_X), "n"(NEPI) : "memory");
    }
  }

  // =========================== softmax partials of this CTA (mppi.py:110-118) ===========================
  trace(-1, 0xA2);                                      // this warp left the horizon loop
  __syncthreads();
  float c = INFINITY;
  if (tid < TM) {
    const float term = quad_full(c_F, s_x, s_const + cl.goalF, nx, p.f_diag, tid);   // mppi.py:79-82, :146-148
    c = s_wgt[tid] + s_cc[tid];                      // helpers' state costs + control warps' action/control costs
    if (p.terminal_mode == 1) c += term;
    else if (valid && (p.k_offset + k_local) == p.K_global - 1) *p.term_out = term;
    if (valid) p.costs[k_local] = c; else c = INFINITY;
    const float m = ampc_warp_min(c);
    if (lane == 0) s_red[warp] = m;
  }
  __syncthreads();
  const float m_cta = fminf(fminf(s_red[0], s_red[1]), fminf(s_red[2], s_red[3]));
  if (tid < TM) {
    const float wgt = valid ? expf(-(c - m_cta) * p.inv_lmda) : 0.f;      // mppi.py:115
    s_wgt[tid] = wgt;
    const float s = ampc_warp_sum(wgt);
    if (lane == 0) s_red[8 + warp] = s;
  }
  __syncthreads();
  float *rec = p.partials + (size_t)blockIdx.x * (2 + HN);
  if (tid == 0) {
    rec[0] = m_cta;
    rec[1] = (s_red[8] + s_red[9]) + (s_red[10] + s_red[11]);
  }
  {                                                                       // mppi.py:117 (per-CTA share)
    // 128 samples of one (step, control) entry are 32 lanes x float4 (the scratch rows are 512-byte aligned);
    // EU entries = EU 16-byte L2 loads in flight per lane
    constexpr int EU = 16;
    const float4 wq = reinterpret_cast<const float4 *>(s_wgt)[lane];
    for (int e0 = warp * EU; e0 < HN; e0 += (NTHR / 32) * EU) {
      float4 ld[EU];
#pragma unroll
      for (int u = 0; u < EU; ++u) {
        const int e = (e0 + u < HN) ? e0 + u : HN - 1;
        ld[u] = __ldcg(reinterpret_cast<const float4 *>(a.epsc + (size_t)e * a.Kc + (size_t)blockIdx.x * TM) + lane);
      }
#pragma unroll
      for (int u = 0; u < EU; ++u) {
        float v = fmaf(wq.x, ld[u].x, fmaf(wq.y, ld[u].y, fmaf(wq.z, ld[u].z, wq.w * ld[u].w)));
        v = ampc_warp_sum(v);
        if (lane == 0 && e0 + u < HN) rec[2 + e0 + u] = v;
      }
    }
  }
  // ---- last CTA to finish merges all partials and applies the update
  trace(-1, 0xA3);                                      // per-CTA softmax record written
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int tk = atomicAdd(p.ticket, 1u);
    s_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    const uint32_t t_merge = (uint32_t)clock();
    // the weight image is dead (every MMA of the pair has completed): its shared memory is the merge's scratch
    if (!ampc_merge_records_tail<32>(p.partials, (int)gridDim.x, 2 + HN, HN, nu, p.inv_lmda, s_act, c_scale, p.act_seq,
                                     p.u_out, p.record_out, reinterpret_cast<float *>(s_w), (int)(a.w_bytes >> 2)))
      ampc_merge_records(p.partials, (int)gridDim.x, 2 + HN, HN, nu, p.inv_lmda, s_act, c_scale, p.act_seq, p.u_out,
                         p.record_out, s_misc);
    if constexpr (TRACE) {
      __syncthreads();
      if (tid == 0)
        printf("# last CTA = %d: merge started %u cycles after its kernel entry and took %u cycles\n", (int)blockIdx.x,
               t_merge - t_entry, (uint32_t)clock() - t_merge);
    }
    if (p.peer_mail != nullptr) ampc_peer_exchange_merge(p, p.record_out, HN, s_act, c_scale, s_misc);
    if (tid == 0) *p.ticket = 0u;
    ampc_publish_host_flag(p);
  }
  // ---- teardown
  trace(-1, 0xA4);                                      // ticket taken / merge done (if this was the last CTA)
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc<CG>(tmem_base);
  if (TRACE && blockIdx.x == 0) {
    if (lane == 0 && trace_n < TRACE_EV) s_trace[warp * TRACE_EV + trace_n++] = ((uint32_t)clock() << 8) | 0xA5u;   // end
    __syncwarp();
    for (int e = lane; e < TRACE_EV; e += 32) {
      const int cnt = __shfl_sync(0xffffffffu, trace_n, 0);
      const uint32_t v = e < cnt ? s_trace[warp * TRACE_EV + e] : 0u;
      a.trace[warp * TRACE_EV + e] = v ? ((unsigned long long)(((v >> 8) - (t_entry & 0xFFFFFFu)) & 0xFFFFFFu) << 8) | (v & 255u) : 0ull;
    }
  }
}


typedef void (*TcKernel)(const AmpcMppiParams, const TcArgs);

}  // namespace ampc_tc

// one getter per instantiation (mppi_tc_inst.cu compiled once per combination)
#define AMPC_TC_DECL(cg, nxp, relu, f16, tr, dz) \
  ampc_tc::TcKernel ampc_tc_kernel_cg##cg##_nxp##nxp##_relu##relu##_f16##f16##_trace##tr##_dz##dz();
#define AMPC_TC_DECL_NXP(cg, relu, f16, dz) \
  AMPC_TC_DECL(cg, 4, relu, f16, 0, dz) AMPC_TC_DECL(cg, 8, relu, f16, 0, dz) AMPC_TC_DECL(cg, 16, relu, f16, 0, dz) \
  AMPC_TC_DECL(cg, 24, relu, f16, 0, dz) AMPC_TC_DECL(cg, 32, relu, f16, 0, dz)
AMPC_TC_DECL_NXP(1, 0, 0, 0) AMPC_TC_DECL_NXP(1, 1, 0, 0) AMPC_TC_DECL_NXP(2, 0, 0, 0) AMPC_TC_DECL_NXP(2, 1, 0, 0)
AMPC_TC_DECL_NXP(1, 0, 1, 0) AMPC_TC_DECL_NXP(1, 1, 1, 0) AMPC_TC_DECL_NXP(2, 0, 1, 0) AMPC_TC_DECL_NXP(2, 1, 1, 0)
// dz mode (ReLU networks only)
AMPC_TC_DECL_NXP(1, 1, 0, 1) AMPC_TC_DECL_NXP(2, 1, 0, 1) AMPC_TC_DECL_NXP(1, 1, 1, 1) AMPC_TC_DECL_NXP(2, 1, 1, 1)
AMPC_TC_DECL(2, 24, 1, 0, 1, 0) AMPC_TC_DECL(2, 24, 1, 0, 1, 1)   // the timeline builds: headline shape only
#undef AMPC_TC_DECL_NXP
#undef AMPC_TC_DECL
