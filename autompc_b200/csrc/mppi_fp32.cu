// MPPI rollout + softmax update, fp32 CUDA-core path (any MLP up to 4x256, any act).
//
// One launch = one MPPI.run (autompc/control/mppi.py:154-168):
//   shift act_sequence (:122-123) -> per step clip / cost / dynamics (:133-144)
//   -> terminal + action cost (:146-150) -> exp-weighted update (:110-118).
//
// Mapping: a CTA owns 32 samples (lane = sample) and has 8 warps; warp w computes
// output neurons [w*NPT, (w+1)*NPT) of every layer for all 32 samples, so a weight
// is one broadcast shared/L1 load per warp and an activation one conflict-free
// load per lane.  Activations live in shared memory as [feature][sample].
// The clipped noise of the CTA's samples is cached in shared memory ([H*nu][32])
// so the weighted control sum needs no second pass over HBM.  Per-CTA softmax
// partials (min, sum w, sum w*eps) are merged by the last CTA to finish
// (log-sum-exp rescaling), which also applies the update -- no second launch.
#include "ampc_common.cuh"

namespace {

constexpr int BM = 32;        // samples per CTA
constexpr int NWARP = 8;
constexpr int NTHR = BM * NWARP;

template <int NPT>
__device__ __forceinline__ void dense_layer(const float *__restrict__ W, const float *__restrict__ B,
                                            const float *hin, float *hout, int Kin, int N, int act,
                                            bool last, int warp, int lane) {
  constexpr int NPAD = NWARP * NPT;
  const float *wb = W + warp * NPT;
  float acc[NPT];
#pragma unroll
  for (int r = 0; r < NPT; ++r) acc[r] = B[warp * NPT + r];
  if constexpr (NPT >= 4) {
#pragma unroll 4
    for (int k = 0; k < Kin; ++k) {
      const float hv = hin[k * BM + lane];
      const float4 *w4 = reinterpret_cast<const float4 *>(wb + (size_t)k * NPAD);
#pragma unroll
      for (int q = 0; q < NPT / 4; ++q) {
        const float4 w = w4[q];
        acc[4 * q + 0] = fmaf(w.x, hv, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(w.y, hv, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(w.z, hv, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(w.w, hv, acc[4 * q + 3]);
      }
    }
  } else {
    // few outputs per thread: split K four ways for instruction-level parallelism
    float p[4][NPT];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int r = 0; r < NPT; ++r) p[u][r] = 0.f;
    int k = 0;
    for (; k + 4 <= Kin; k += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float hv = hin[(k + u) * BM + lane];
#pragma unroll
        for (int r = 0; r < NPT; ++r) p[u][r] = fmaf(wb[(size_t)(k + u) * NPAD + r], hv, p[u][r]);
      }
    }
    for (; k < Kin; ++k) {
      const float hv = hin[k * BM + lane];
#pragma unroll
      for (int r = 0; r < NPT; ++r) p[0][r] = fmaf(wb[(size_t)k * NPAD + r], hv, p[0][r]);
    }
#pragma unroll
    for (int r = 0; r < NPT; ++r) acc[r] += (p[0][r] + p[1][r]) + (p[2][r] + p[3][r]);
  }
#pragma unroll
  for (int r = 0; r < NPT; ++r) {
    const int j = warp * NPT + r;
    if (j < N) hout[j * BM + lane] = last ? acc[r] : ampc_act<float>(act, acc[r]);
  }
}

__device__ __forceinline__ void dense_dispatch(int npt, const float *W, const float *B, const float *hin,
                                               float *hout, int Kin, int N, int act, bool last, int warp,
                                               int lane) {
  switch (npt) {
    case 1: dense_layer<1>(W, B, hin, hout, Kin, N, act, last, warp, lane); break;
    case 2: dense_layer<2>(W, B, hin, hout, Kin, N, act, last, warp, lane); break;
    case 4: dense_layer<4>(W, B, hin, hout, Kin, N, act, last, warp, lane); break;
    case 8: dense_layer<8>(W, B, hin, hout, Kin, N, act, last, warp, lane); break;
    case 16: dense_layer<16>(W, B, hin, hout, Kin, N, act, last, warp, lane); break;
    default: dense_layer<32>(W, B, hin, hout, Kin, N, act, last, warp, lane); break;
  }
}

// quadratic form rows i = warp, warp+8, ... of  d^T M d  for this lane's sample
__device__ __forceinline__ float quad_rows(const float *M, const float *v, const float *off, int n,
                                           bool diag, int warp, int lane) {
  float c = 0.f;
  for (int i = warp; i < n; i += NWARP) {
    const float di = v[i * BM + lane] - (off ? off[i] : 0.f);
    if (diag) {
      c = fmaf(M[i * n + i] * di, di, c);
    } else {
      float row = 0.f;
      for (int j = 0; j < n; ++j) row = fmaf(M[i * n + j], v[j * BM + lane] - (off ? off[j] : 0.f), row);
      c = fmaf(di, row, c);
    }
  }
  return c;
}

template <bool RES>
__global__ void __launch_bounds__(NTHR) mppi_rollout_fp32_kernel(const AmpcMppiParams p) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nx = p.nx, nu = p.nu, H = p.H, HN = H * nu;
  const AmpcConstLayout cl(nx, nu);

  float *s_const = smem;
  float *s_act = s_const + cl.total;                 // shifted act_sequence (H*nu)
  float *s_x = s_act + ((HN + 3) & ~3);              // state  [nx][32]
  float *s_u = s_x + nx * BM;                        // scaled control [nu][32]
  float *s_h0 = s_u + nu * BM;                       // activations ping
  float *s_h1 = s_h0 + p.max_width * BM;             // activations pong
  float *s_eps = s_h1 + p.max_width * BM;            // clipped noise [H*nu][32]
  float *s_red = s_eps + HN * BM;                    // 2 * 8 * 32 reduction scratch
  float *s_misc = s_red + 2 * NWARP * BM;            // 64 + AMPC_MERGE_CACHE
  float *s_w = s_misc + 64 + AMPC_MERGE_CACHE;       // resident weights (RES)
  __shared__ int s_last;

  for (int i = tid; i < cl.total; i += NTHR) s_const[i] = p.consts[i];
  for (int e = tid; e < HN; e += NTHR) {             // mppi.py:122-123
    const int i = e / nu, j = e - i * nu;
    const int src = (i + 1 < H) ? i + 1 : H - 1;
    s_act[e] = p.act_seq[src * nu + j];
  }
  if (RES) {
    const float4 *g4 = reinterpret_cast<const float4 *>(p.wpack);
    float4 *s4 = reinterpret_cast<float4 *>(s_w);
    for (int i = tid; i < p.wpack_floats / 4; i += NTHR) s4[i] = g4[i];
  }
  for (int j = warp; j < nx; j += NWARP) s_x[j * BM + lane] = (p.x0_inline && j < 32) ? p.x0_val[j] : p.x0[j];   // mppi.py:129-130
  __syncthreads();

  const float *wbase = RES ? s_w : p.wpack;
  const float *c_mean = s_const + cl.xu_mean, *c_inv = s_const + cl.xu_inv;
  const float *c_dym = s_const + cl.dy_mean, *c_dys = s_const + cl.dy_std;
  const float *c_goal = s_const + cl.goal, *c_Q = s_const + cl.Q, *c_R = s_const + cl.R, *c_F = s_const + cl.F;
  const float *c_lo = s_const + cl.lo, *c_hi = s_const + cl.hi, *c_scale = s_const + cl.scale;

  const int k_local = blockIdx.x * BM + lane;
  const bool valid = k_local < p.K;
  const uint32_t kg = (uint32_t)(p.k_offset + k_local);
  const int nblk = (nu + 3) >> 2;
  float cost_acc = 0.f;   // this thread's share of the sample's running cost

  for (int i = 0; i < H; ++i) {
    // ---- controls: noise, clip, write-back (mppi.py:134-139), action cost (:143)
    for (int blk = warp; blk < nblk; blk += NWARP) {
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
          const float a = fminf(c_hi[j], fmaxf(c_lo[j], n4[q] + a0));
          const float e = a - a0;
          s_eps[(i * nu + j) * BM + lane] = e;
          s_u[j * BM + lane] = a * c_scale[j];
          cost_acc = fmaf(p.lam_over_sigma * a, e, cost_acc);
        }
      }
    }
    __syncthreads();
    // ---- stage cost at the pre-step state (mppi.py:142; cost.py:79-81, :131-132) + z-score (mlp.py:20-24)
    cost_acc += quad_rows(c_Q, s_x, c_goal, nx, p.q_diag, warp, lane);
    cost_acc += quad_rows(c_R, s_u, nullptr, nu, p.r_diag, warp, lane);
    for (int b = warp; b < p.n_box; b += NWARP) {          // threshold terms (thresh_cost.py:27-32, :73-77)
      const float *bx = p.box + (size_t)b * (2 * nx + 1);
      bool out = false;
      for (int j = 0; j < nx; ++j) {
        const float x = s_x[j * BM + lane];
        out = out || (x < __ldg(bx + j)) || (x > __ldg(bx + nx + j));
      }
      if (out) cost_acc += __ldg(bx + 2 * nx);
    }
    for (int j = warp; j < nx + nu; j += NWARP) {
      const float v = (j < nx) ? s_x[j * BM + lane] : s_u[(j - nx) * BM + lane];
      s_h0[j * BM + lane] = (v - c_mean[j]) * c_inv[j];
    }
    __syncthreads();
    // ---- MLP (mlp.py:55-59)
    float *hin = s_h0, *hout = s_h1;
    for (int l = 0; l < p.n_layers; ++l) {
      dense_dispatch(p.npt[l], wbase + p.woff[l], wbase + p.boff[l], hin, hout, p.dims[l], p.dims[l + 1],
                     p.act, l == p.n_layers - 1, warp, lane);
      __syncthreads();
      float *t = hin; hin = hout; hout = t;
    }
    // ---- un-z-score + integrate (mlp.py:235-236); hin now holds the net output
    for (int j = warp; j < nx; j += NWARP)
      s_x[j * BM + lane] += fmaf(hin[j * BM + lane], c_dys[j], c_dym[j]);
    __syncthreads();
  }

  // ---- terminal cost (mppi.py:79-82, :146-148)
  const float term_acc = quad_rows(c_F, s_x, s_const + cl.goalF, nx, p.f_diag, warp, lane);
  s_red[warp * BM + lane] = cost_acc;
  s_red[(NWARP + warp) * BM + lane] = term_acc;
  __syncthreads();
  float *s_wgt = s_h0;   // per-sample softmax numerators (activations no longer needed)
  if (warp == 0) {
    float c = 0.f, t = 0.f;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) { c += s_red[w * BM + lane]; t += s_red[(NWARP + w) * BM + lane]; }
    if (p.terminal_mode == 1) c += t;
    else if (valid && (int)kg == p.K_global - 1) *p.term_out = t;
    if (valid) p.costs[k_local] = c; else c = INFINITY;
    const float m = ampc_warp_min(c);
    const float wgt = valid ? expf(-(c - m) * p.inv_lmda) : 0.f;     // mppi.py:115
    const float s = ampc_warp_sum(wgt);
    s_wgt[lane] = wgt;
    if (lane == 0) {
      float *rec = p.partials + (size_t)blockIdx.x * (2 + HN);
      rec[0] = m;
      rec[1] = s;
    }
  }
  __syncthreads();
  {
    float *rec = p.partials + (size_t)blockIdx.x * (2 + HN) + 2;
    const float wgt = s_wgt[lane];
    for (int e = warp; e < HN; e += NWARP) {                          // mppi.py:117 (per-CTA share)
      const float v = ampc_warp_sum(wgt * s_eps[e * BM + lane]);
      if (lane == 0) rec[e] = v;
    }
  }
  // ---- last CTA to finish merges all partials and applies the update
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int t = atomicAdd(p.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    ampc_merge_records(p.partials, (int)gridDim.x, 2 + HN, HN, nu, p.inv_lmda, s_act, c_scale, p.act_seq, p.u_out,
                       p.record_out, s_misc);
    if (p.peer_mail != nullptr) ampc_peer_exchange_merge(p, p.record_out, HN, s_act, c_scale, s_misc);
    if (tid == 0) *p.ticket = 0u;
    ampc_publish_host_flag(p);
  }
}

}  // namespace

size_t ampc_mppi_fp32_smem_bytes(const AmpcMppiParams &p, bool resident) {
  const AmpcConstLayout cl(p.nx, p.nu);
  const int HN = p.H * p.nu;
  size_t f = cl.total + ((HN + 3) & ~3) + (size_t)p.nx * BM + (size_t)p.nu * BM + 2 * (size_t)p.max_width * BM +
             (size_t)HN * BM + 2 * NWARP * BM + 64 + AMPC_MERGE_CACHE;
  if (resident) f += p.wpack_floats;
  return f * sizeof(float);
}

int ampc_mppi_fp32_grid(const AmpcMppiParams &p) { return (p.K + BM - 1) / BM; }

int ampc_mppi_fp32_configure(const AmpcMppiParams &p, bool *resident_out, size_t *smem_out) {
  int dev = 0, max_optin = 0;
  AMPC_CUDA_CHECK(cudaGetDevice(&dev));
  AMPC_CUDA_CHECK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t cap = (size_t)max_optin - 1024;
  size_t need = ampc_mppi_fp32_smem_bytes(p, true);
  bool res = need <= cap;
  if (!res) need = ampc_mppi_fp32_smem_bytes(p, false);
  AMPC_REQUIRE(need <= cap, AMPC_ERR_UNSUPPORTED,
               "fp32 MPPI kernel needs %zu B shared memory (H*nu=%d too large), limit %zu", need,
               p.H * p.nu, cap);
  if (res)
    AMPC_CUDA_CHECK(ampc_raise_smem_limit((const void *)mppi_rollout_fp32_kernel<true>, need));
  else
    AMPC_CUDA_CHECK(ampc_raise_smem_limit((const void *)mppi_rollout_fp32_kernel<false>, need));
  *resident_out = res;
  *smem_out = need;
  return AMPC_OK;
}

int ampc_mppi_fp32_launch(const AmpcMppiParams &p, bool resident, size_t smem, cudaStream_t stream) {
  const int grid = ampc_mppi_fp32_grid(p);
  if (resident)
    mppi_rollout_fp32_kernel<true><<<grid, NTHR, smem, stream>>>(p);
  else
    mppi_rollout_fp32_kernel<false><<<grid, NTHR, smem, stream>>>(p);
  ampc_count_launch();
  AMPC_CUDA_CHECK(cudaGetLastError());
  return AMPC_OK;
}
