// Shared device/host helpers for libampc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/ampc_b200.h"

// ------------------------------------------------------------------ errors ---
void ampc_set_error(const char *fmt, ...);
void ampc_count_launch(int n = 1);
// cudaFuncAttributeMaxDynamicSharedMemorySize is per FUNCTION, not per handle: handles with different shared-memory
// needs share the kernels, so the limit is only ever raised (process-wide maximum per function and device).
cudaError_t ampc_raise_smem_limit(const void *func, size_t bytes);

#define AMPC_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ampc_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                     cudaGetErrorString(_e));                                              \
      return AMPC_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define AMPC_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) {                    \
      ampc_set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

// ----------------------------------------------------------------- Philox ---
// Philox4x32-10 (Salmon et al. 2011).  Counter = (global sample, step | block<<16,
// solve counter lo, hi); key = (seed lo, seed hi).  One call yields four N(0,1)
// draws (two Box-Muller pairs) for control dims [4*block, 4*block+4).
__host__ __device__ __forceinline__ void ampc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                            uint32_t c3, uint32_t k0, uint32_t k1,
                                                            uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), hi1 = __umulhi(0xCD9E8D57u, c2);
#else
    const uint32_t hi0 = (uint32_t)(((uint64_t)0xD2511F53u * c0) >> 32);
    const uint32_t hi1 = (uint32_t)(((uint64_t)0xCD9E8D57u * c2) >> 32);
#endif
    const uint32_t lo0 = 0xD2511F53u * c0, lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Four unit normals for (sample kg, step h, 4-wide control block blk).
__device__ __forceinline__ void ampc_normal4(uint64_t seed, uint64_t ctr, uint32_t kg, uint32_t h,
                                             uint32_t blk, float n[4]) {
  uint32_t r[4];
  ampc_philox4x32_10(kg, h | (blk << 16), (uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed,
                     (uint32_t)(seed >> 32), r);
  const float inv24 = 5.9604644775390625e-8f;  // 2^-24
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float u1 = (float)((r[2 * q] >> 8) + 1u) * inv24;  // (0,1]
    const float u2 = (float)(r[2 * q + 1] >> 8) * inv24;     // [0,1)
#ifdef __CUDA_ARCH__
    // SFU path: lg2 / sin / cos approximations (abs error ~1e-6 on the unit normal, checked against the
    // float64 restatement in oracle/philox.py); the angle is reduced to [-pi, pi) first.
    const float rad = sqrtf(-1.3862943611198906f * __log2f(u1));
    const float ang = 6.283185307179586f * (u2 - 0.5f);
    n[2 * q] = -rad * __cosf(ang);
    n[2 * q + 1] = -rad * __sinf(ang);
#else
    const float rad = sqrtf(-2.0f * logf(u1));
    n[2 * q] = rad * cosf(6.283185307179586f * u2);
    n[2 * q + 1] = rad * sinf(6.283185307179586f * u2);
#endif
  }
}

// ------------------------------------------------------------ activations ---
template <typename T>
__device__ __forceinline__ T ampc_act(int act, T y) {
  switch (act) {
    case AMPC_ACT_RELU: return y > T(0) ? y : T(0);
    case AMPC_ACT_TANH: return tanh(y);
    case AMPC_ACT_SIGMOID: return T(1) / (T(1) + exp(-y));
    default: {  // SELU, torch.nn.SELU constants
      const T alpha = T(1.6732632423543772848170429916717), scale = T(1.0507009873554804934193349852946);
      return scale * (y > T(0) ? y : alpha * expm1(y));
    }
  }
}
template <typename T>
__device__ __forceinline__ T ampc_act_grad(int act, T y) {
  switch (act) {
    case AMPC_ACT_RELU: return y > T(0) ? T(1) : T(0);
    case AMPC_ACT_TANH: { T t = tanh(y); return T(1) - t * t; }
    case AMPC_ACT_SIGMOID: { T s = T(1) / (T(1) + exp(-y)); return s * (T(1) - s); }
    default: {
      const T alpha = T(1.6732632423543772848170429916717), scale = T(1.0507009873554804934193349852946);
      return scale * (y > T(0) ? T(1) : alpha * exp(y));
    }
  }
}

// --------------------------------------------------------------- reductions ---
__device__ __forceinline__ float ampc_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float ampc_warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ----------------------------------------------------- MPPI kernel params ---
// Layout of the fp32 constant block both rollout kernels stage into shared memory.
struct AmpcConstLayout {
  int xu_mean, xu_inv, dy_mean, dy_std, goal, goalF, Q, R, F, lo, hi, scale, total;
  __host__ __device__ AmpcConstLayout(int nx, int nu) {
    int o = 0;
    xu_mean = o; o += nx + nu;
    xu_inv = o; o += nx + nu;
    dy_mean = o; o += nx;
    dy_std = o; o += nx;
    goal = o; o += nx;
    goalF = o; o += nx;   // goal of the terminal term (== goal unless a SumCost was folded)
    Q = o; o += nx * nx;
    R = o; o += nu * nu;
    F = o; o += nx * nx;
    lo = o; o += nu;
    hi = o; o += nu;
    scale = o; o += nu;
    total = (o + 3) & ~3;
  }
};

struct AmpcMppiParams {
  int K, H, nx, nu;
  int k_offset, K_global;
  int terminal_mode, q_diag, f_diag, r_diag, act;
  int n_layers;
  int dims[AMPC_MAX_LAYERS + 1];
  int npt[AMPC_MAX_LAYERS];   // fp32 kernel: outputs per thread of each layer
  int woff[AMPC_MAX_LAYERS];  // float offsets into wpack
  int boff[AMPC_MAX_LAYERS];
  int wpack_floats;
  int max_width;              // widest activation (incl. input)
  float inv_lmda, sqrt_sigma, lam_over_sigma;
  uint64_t seed, ctr;
  const float *wpack;     // packed weights + biases (layout is kernel-specific)
  const float *consts;    // AmpcConstLayout
  const float *x0;        // (nx,)
  const float *eps;       // NULL or (H,K,nu) unclipped noise
  float *act_seq;         // (H,nu) in/out
  float *costs;           // (K,)
  float *term_out;        // scalar (terminal_mode 0)
  float *partials;        // (gridDim.x, 2 + H*nu)
  unsigned int *ticket;   // grid completion counter
  float *record_out;      // NULL: update act_seq in-kernel;  else write [m, s, W] here
  // NVLink peer-memory exchange (multi-GPU, fused into the rollout kernel's tail); peer_mail == NULL: off.
  float *const *peer_mail;  // device array of `world` mailbox base pointers (own mailbox at index `rank`)
  int world, rank;
  unsigned int seq;         // solve sequence number (> 0), identical on all ranks; selects the mailbox slot
  float *u_out;           // (nu,)
  // threshold stage costs (thresh_cost.py): n_box terms, each [lo (nx) | hi (nx) | weight] in `box`
  int n_box;
  const float *box;
  // host-buffer entry point: the observation rides in the kernel parameters (no H2D copy on the stream) ...
  int x0_inline;          // != 0: use x0_val instead of x0
  float x0_val[32];       // nx <= 32 on this path
  // ... and the last CTA publishes "control written" in mapped pinned host memory, so that the host can return as
  // soon as the result has landed instead of waiting for the stream to drain (null: off)
  unsigned int *host_flag;
  unsigned int host_seq;
};

// Called by every thread of the CTA that finished the solve, after the control has been written.
__device__ __forceinline__ void ampc_publish_host_flag(const AmpcMppiParams &p) {
  if (p.host_flag == nullptr) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned int *>(p.host_flag) = p.host_seq;
  }
}

// Merge softmax partial records [m, s, W(HN)] (block level or rank level) and either
// apply the update (mppi.py:115-118) or emit the merged record.  Called by one CTA.
// `s_act_shift` = shifted action sequence (mppi.py:122-123) in shared memory.
// s_scratch: >= 64 + AMPC_MERGE_CACHE floats of shared memory.
#define AMPC_MERGE_CACHE 1024
__device__ __forceinline__ void ampc_merge_records(const float *recs, int n_recs, int rec_stride, int HN,
                                                   int nu, float inv_lmda, const float *s_act_shift,
                                                   const float *scale, float *act_seq, float *u_out,
                                                   float *record_out, float *s_scratch) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  float *s_rs = s_scratch + 64;  // per-record rescale exp(-(m_b - m)/lmda), cached when it fits
  const bool cached = n_recs <= AMPC_MERGE_CACHE;
  float m = INFINITY;
  for (int b = tid; b < n_recs; b += nthr) m = fminf(m, __ldcg(recs + (size_t)b * rec_stride));
  m = ampc_warp_min(m);
  if (lane == 0) s_scratch[warp] = m;
  __syncthreads();
  m = s_scratch[0];
  for (int w = 1; w < nwarp; ++w) m = fminf(m, s_scratch[w]);
  float s = 0.f;
  for (int b = tid; b < n_recs; b += nthr) {
    const float mb = __ldcg(recs + (size_t)b * rec_stride);
    const float rs = expf(-(mb - m) * inv_lmda);
    if (cached) s_rs[b] = rs;
    s += __ldcg(recs + (size_t)b * rec_stride + 1) * rs;
  }
  s = ampc_warp_sum(s);
  if (lane == 0) s_scratch[32 + warp] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < nwarp; ++w) s += s_scratch[32 + w];
  // W[e] = sum_b rec[b][2+e] * rs[b].  Fast path (records in L2, everything 8-byte aligned, scratch large enough):
  // thread (g, pair) sums every G-th record for two neighbouring entries with MU float2 loads in flight, the G
  // partial sums meet in shared memory.  The one-thread-per-entry loop below is the general path.
  const int E2 = HN >> 1;
  int G = E2 > 0 ? nthr / E2 : 0;
  if (G > 8) G = 8;
  if (G > n_recs) G = n_recs;
  const bool fast = cached && G >= 1 && (HN & 1) == 0 && (rec_stride & 1) == 0 && n_recs + G * HN <= AMPC_MERGE_CACHE;
  if (fast) {
    float *s_part = s_rs + n_recs;                   // [G][HN]
    const int g = tid / E2, pe = tid - g * E2;
    if (g < G) {
      constexpr int MU = 16;
      float2 acc = make_float2(0.f, 0.f);
      const float *base = recs + 2 + 2 * pe;
      for (int b0 = g; b0 < n_recs; b0 += G * MU) {
        float2 v[MU];
#pragma unroll
        for (int u = 0; u < MU; ++u) {
          const int b = b0 + u * G;
          v[u] = b < n_recs ? __ldcg(reinterpret_cast<const float2 *>(base + (size_t)b * rec_stride)) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < MU; ++u) {
          const int b = b0 + u * G;
          const float r = b < n_recs ? s_rs[b] : 0.f;
          acc.x = fmaf(v[u].x, r, acc.x);
          acc.y = fmaf(v[u].y, r, acc.y);
        }
      }
      s_part[g * HN + 2 * pe] = acc.x;
      s_part[g * HN + 2 * pe + 1] = acc.y;
    }
    __syncthreads();
    for (int e = tid; e < HN; e += nthr) {
      float acc = 0.f;
      for (int gg = 0; gg < G; ++gg) acc += s_part[gg * HN + e];
      if (record_out) {
        record_out[2 + e] = acc;
      } else {
        const float v = s_act_shift[e] + acc / s;
        act_seq[e] = v;
        if (e < nu) u_out[e] = v * scale[e];
      }
    }
  } else
  for (int e = tid; e < HN; e += nthr) {
    float acc = 0.f;
    if (cached) {
      for (int b = 0; b < n_recs; ++b)
        acc = fmaf(__ldcg(recs + (size_t)b * rec_stride + 2 + e), s_rs[b], acc);
    } else {
      for (int b = 0; b < n_recs; ++b) {
        const float mb = __ldcg(recs + (size_t)b * rec_stride);
        acc = fmaf(__ldcg(recs + (size_t)b * rec_stride + 2 + e), expf(-(mb - m) * inv_lmda), acc);
      }
    }
    if (record_out) {
      record_out[2 + e] = acc;
    } else {
      const float v = s_act_shift[e] + acc / s;
      act_seq[e] = v;
      if (e < nu) u_out[e] = v * scale[e];
    }
  }
  if (record_out && tid == 0) {
    record_out[0] = m;
    record_out[1] = s;
  }
}

// The same merge for the in-kernel tail of the tensor-core kernel, where it is on the critical path of every solve
// (one CTA, ~1 % of the kernel per 5 k cycles) and a large scratch is free (the weight image is dead by then):
//   * (m_b, s_b) of every record are fetched ONCE (one float2 each) and kept in shared memory -- the generic routine
//     walks the records three times, each walk a dependent L2 round trip of lines other SMs have just written;
//   * the first MU records of every thread's share of W are requested BEFORE that, so their latency hides under the
//     min / sum passes; MU = 32 loads in flight per thread leave one more round trip for 128 records.
// Needs HN and rec_stride even and 64 + 3 n_recs + G HN floats of scratch; returns false (nothing done) otherwise.
template <int MU>
__device__ __forceinline__ bool ampc_merge_records_tail(const float *recs, int n_recs, int rec_stride, int HN, int nu,
                                                        float inv_lmda, const float *s_act_shift, const float *scale,
                                                        float *act_seq, float *u_out, float *record_out, float *s_scratch,
                                                        int scratch_floats) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  const int E2 = HN >> 1;
  int G = E2 > 0 ? nthr / E2 : 0;
  if (G > 8) G = 8;
  if (G > n_recs) G = n_recs;
  if (G < 1 || (HN & 1) || (rec_stride & 1) || 64 + 3 * n_recs + G * HN > scratch_floats) return false;   // uniform
  float *s_m = s_scratch + 64, *s_s = s_m + n_recs, *s_rs = s_s + n_recs, *s_part = s_rs + n_recs;        // [G][HN]
  const int g = tid / E2, pe = tid - g * E2;
  const bool active = g < G;
  const float *base = recs + 2 + 2 * pe;
  float2 v[MU];
  if (active) {
#pragma unroll
    for (int u = 0; u < MU; ++u) {
      const int b = g + u * G;
      v[u] = b < n_recs ? __ldcg(reinterpret_cast<const float2 *>(base + (size_t)b * rec_stride)) : make_float2(0.f, 0.f);
    }
  }
  float m = INFINITY;
  for (int b = tid; b < n_recs; b += nthr) {
    const float2 ms = __ldcg(reinterpret_cast<const float2 *>(recs + (size_t)b * rec_stride));
    s_m[b] = ms.x;
    s_s[b] = ms.y;
    m = fminf(m, ms.x);
  }
  m = ampc_warp_min(m);
  if (lane == 0) s_scratch[warp] = m;
  __syncthreads();
  m = s_scratch[0];
  for (int w = 1; w < nwarp; ++w) m = fminf(m, s_scratch[w]);
  float s = 0.f;
  for (int b = tid; b < n_recs; b += nthr) {
    const float rs = expf(-(s_m[b] - m) * inv_lmda);
    s_rs[b] = rs;
    s += s_s[b] * rs;
  }
  s = ampc_warp_sum(s);
  if (lane == 0) s_scratch[32 + warp] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < nwarp; ++w) s += s_scratch[32 + w];
  if (active) {
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < MU; ++u) {
      const int b = g + u * G;
      const float r = b < n_recs ? s_rs[b] : 0.f;
      acc.x = fmaf(v[u].x, r, acc.x);
      acc.y = fmaf(v[u].y, r, acc.y);
    }
    for (int b0 = g + MU * G; b0 < n_recs; b0 += G * MU) {
#pragma unroll
      for (int u = 0; u < MU; ++u) {
        const int b = b0 + u * G;
        v[u] = b < n_recs ? __ldcg(reinterpret_cast<const float2 *>(base + (size_t)b * rec_stride)) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < MU; ++u) {
        const int b = b0 + u * G;
        const float r = b < n_recs ? s_rs[b] : 0.f;
        acc.x = fmaf(v[u].x, r, acc.x);
        acc.y = fmaf(v[u].y, r, acc.y);
      }
    }
    s_part[g * HN + 2 * pe] = acc.x;
    s_part[g * HN + 2 * pe + 1] = acc.y;
  }
  __syncthreads();
  for (int e = tid; e < HN; e += nthr) {
    float acc = 0.f;
    for (int gg = 0; gg < G; ++gg) acc += s_part[gg * HN + e];
    if (record_out) {
      record_out[2 + e] = acc;
    } else {
      const float val = s_act_shift[e] + acc / s;
      act_seq[e] = val;
      if (e < nu) u_out[e] = val * scale[e];
    }
  }
  if (record_out && tid == 0) {
    record_out[0] = m;
    record_out[1] = s;
  }
  return true;
}

// ------------------------------------------------------- NVLink peer exchange ---
// Mailbox of a rank (in its own HBM, mapped into every peer): 2 slots x [ world records of `rec` floats ]
// followed by 2 x world 32-bit flags.  Called by ONE CTA per GPU (the last to finish) once the shard's record
// [m, s, W(HN)] is in `rec_local` (global): store it into slot seq&1 / column `rank` of EVERY rank's mailbox
// over NVLink, publish with a system-scope fence + flag = seq, wait for the `world` flags of the own mailbox,
// then merge the records in rank order (identical arithmetic on all ranks) and apply mppi.py:115-118.
__device__ __forceinline__ size_t ampc_mail_floats(int world, int rec) { return (size_t)2 * world * rec + 2 * world; }

__device__ __forceinline__ void ampc_peer_exchange_merge(const AmpcMppiParams &p, const float *rec_local, int HN,
                                                         const float *s_act_shift, const float *scale,
                                                         float *s_scratch) {
  const int rec = 2 + HN, world = p.world, slot = (int)(p.seq & 1u);
  const int tid = threadIdx.x, nthr = blockDim.x;
  __syncthreads();
  __threadfence();                                       // rec_local was written by this CTA's other threads
  for (int r = 0; r < world; ++r) {
    float *dst = p.peer_mail[r] + ((size_t)slot * world + p.rank) * rec;
    for (int e = tid; e < rec; e += nthr) dst[e] = __ldcg(rec_local + e);
  }
  __threadfence_system();
  __syncthreads();
  if (tid < world) {
    unsigned int *flags = reinterpret_cast<unsigned int *>(p.peer_mail[tid] + (size_t)2 * world * rec);
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags + slot * world + p.rank), "r"(p.seq) : "memory");
  }
  if (tid < world) {
    const unsigned int *flags = reinterpret_cast<const unsigned int *>(p.peer_mail[p.rank] + (size_t)2 * world * rec);
    unsigned int v = 0, spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + slot * world + tid) : "memory");
      if (++spins > (1u << 26)) {
        printf("ampc: peer exchange timeout (rank %d waiting for rank %d, seq %u, saw %u)\n", p.rank, tid, p.seq, v);
        __trap();
      }
    } while (v != p.seq);
  }
  __syncthreads();
  __threadfence_system();
  ampc_merge_records(p.peer_mail[p.rank] + (size_t)slot * world * rec, world, rec, HN, p.nu, p.inv_lmda, s_act_shift,
                     scale, p.act_seq, p.u_out, nullptr, s_scratch);
}
