// Iterative LQR solve on the device, float64, one persistent CTA per solve.
//
// Replaces autompc.control.ilqr.IterativeLQR.compute_ilqr_default
// (autompc/control/ilqr.py:100-265) for MLP dynamics + QuadCost:
//   init rollout (:141-149)  ->  up to max_iter x [ backward Riccati (:159-187),
//   batched ls_max_iter-alpha line search (:197-225), Jacobian refresh (:226-234),
//   stopping rule (:235-261) ].
// The problem is tiny and strictly sequential in H (state recursion forward, value recursion backward), so the whole
// solve is ONE launch: no host round trips between the ~50 x (H Riccati steps + H line-search steps) stages the
// reference walks in Python.  float64 keeps the integer outputs (adopted line-search index per iteration, iteration
// count, converged) equal to the reference's.
//
// Work decomposition inside the CTA (20 warps):
//   * backward pass: ONE warp, warp barriers only -- the matrices are (nx+nu)^2; (row, column) of every element comes
//     from small index tables instead of integer divisions;
//   * line search: all step sizes batched per horizon step like the reference (ilqr.py:197-205) by 8 warps = 64 output
//     slots x 4 K-quarters; a weight is read from shared memory once for all alphas (ls_rollouts);
//   * Jacobian refresh (mlp.py:281-305 in closed form): one warp per horizon step, forward-mode propagation of the
//     [width x (nx+nu)] panel with a strip of two rows held in registers (jac_batch).
// Data placement (template RES): when the network (row-major, rows padded to an odd stride so that lane j reading row j
// is bank-conflict free), the trajectories / gains / Jacobians / line-search rollouts and the per-warp Jacobian scratch
// all fit in shared memory (the cartpole problem: 40 + 37 + 123 KB) they live there and every pointer is derived from
// the shared array, so the compiler emits LDS/STS with immediate offsets; otherwise the same code runs on global
// scratch.  What the profiles said on the way (profiles/r02_ilqr_*): the round-1 kernel and the first round-2 cut spent
// 111 M warp instructions per solve of which 10 M were DFMA -- 64-bit generic address arithmetic; with that gone, rolling
// every alpha out on its own was bound by the shared-memory pipe (the whole network per alpha per step), and the
// one-warp Riccati recursion by its instruction count (a lone warp retires a dependent instruction every ~10 cycles).
#include <algorithm>
#include <vector>

#include "ampc_common.cuh"

namespace {

constexpr int NT = 512;            // 16 warps (<= 128 registers per thread)
constexpr int NWARPS = NT / 32;
constexpr int JAC_WARPS = 16;      // warps that take part in the Jacobian refresh (bounds its shared-memory scratch)
constexpr int LS_WARPS = 16;       // warps of the line-search phase (4 or 8 of them run the main loops)
constexpr int MAX_NU = 16;
constexpr int MAX_LS = 20;
constexpr int MAXL = AMPC_MAX_LAYERS;

struct IlqrNet {                   // offsets in doubles from the base of the network blob
  int n_layers, act, max_width, total;
  int dims[MAXL + 1];
  int woff[MAXL], wstride[MAXL], boff[MAXL];   // W_l row-major (out, in) with an odd row stride: lane j reading row j is conflict free
  int w0s;                                     // W_0[j][c] / xu_std[c], row stride dims[0]: first Jacobian panel (mlp.py:298)
  int xu_mean, xu_std, dy_mean, dy_std;
};

struct IlqrParams {
  IlqrNet net;
  const double *net_blob;          // global copy of the blob
  int H, nx, nu, bounded, max_iter, ls_max_iter;
  double dt, ls_discount, ls_cost_threshold, u_threshold;
  const double *cst;               // global: Q | R | F | goal | goalF | umin | umax | alphas
  const double *x0, *uguess;       // global (uguess may be null)
  double *traj;                    // global scratch for the trajectory block when it is not shared-memory resident
  double *jac_work;                // global per-warp scratch of the Jacobian refresh (ditto)
  double *states, *ctrls, *Ks, *ks;   // outputs (global)
  int *info, *alpha_idx;
  // line search on FP64 tensor-core fragments (ls_rollouts_mma): on when the fragment-ordered weight image fits
  int jac_mma, jac_SP, ls_wf_total;   // Jacobian refresh on the tensor-core path: on, panel row stride, doubles of the Wf image
  int ls_profile;                  // != 0: ls_rollouts_mma reads the clock at its phase boundaries (AMPC_ILQR_LS_PROFILE)
  int ls_mma, ls_S, ls_mwp, ls_doubles;   // activation row stride (doubles), padded activation rows, phase scratch in doubles
  int ls_wf[MAXL], ls_kp[MAXL], ls_mt[MAXL];   // per layer: offset of its image inside the Wf region, padded K, 8-row tiles
  unsigned long long *prof;        // [0..5] cycles of thread 0 in: setup+init rollout, backward passes, line-search rollouts,
                                   // objective + acceptance, Jacobian refreshes, copy-out; [6] = total, [7] = iterations
};

// offsets inside the trajectory block
struct TrajLayout {
  size_t states, ctrls, Ks, ks, Jacs, ls_states, ls_ctrls, step_cost, total;
  __host__ __device__ TrajLayout(int H, int nx, int nu, int LS) {
    const size_t n = nx + nu;
    size_t o = 0;
    auto take = [&](size_t c) { size_t r = o; o += (c + 1) & ~(size_t)1; return r; };
    states = take((size_t)(H + 1) * nx); ctrls = take((size_t)H * nu); Ks = take((size_t)H * nu * nx);
    ks = take((size_t)H * nu); Jacs = take((size_t)H * nx * n); ls_states = take((size_t)LS * (H + 1) * nx);
    ls_ctrls = take((size_t)LS * H * nu); step_cost = take((size_t)(LS + 1) * (H + 1));
    total = o;
  }
};

__host__ __device__ inline int jac_cols(int n) { return (n + 1) & ~1; }           // panel row stride: even, rows 16-byte aligned
__host__ __device__ inline size_t jac_per_warp(int mw, int n) {
  return 4 * (size_t)((mw + 1) & ~1) + 2 * (size_t)mw * jac_cols(n);
}
__host__ __device__ inline int ls_cols(int LS) { return LS <= 10 ? 10 : 20; }   // activation columns of ls_rollouts (AG x LAP)
__host__ __device__ inline size_t ls_scratch(int mw, int LS) { return (size_t)6 * mw * ls_cols(LS); }   // hT x 2 + partials x 4
__host__ __device__ inline size_t cst_doubles(int nx, int nu, int LS) {
  return 2 * (size_t)nx * nx + (size_t)nu * nu + 2 * (size_t)nx + 2 * (size_t)nu + LS;
}

__device__ __forceinline__ double block_sum(double v, double *s_red, int tid) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < NWARPS; ++w) r += s_red[w];
  __syncthreads();
  return r;
}

// x^T M x with the reference's evaluation shape (obst.T @ M @ obst, cost.py:81)
__device__ __forceinline__ double quad_form(const double *M, const double *x, const double *off, int n) {
  double tot = 0.0;
  for (int j = 0; j < n; ++j) {
    double col = 0.0;
    for (int i = 0; i < n; ++i) col += (x[i] - (off ? off[i] : 0.0)) * M[i * n + j];
    tot += col * (x[j] - (off ? off[j] : 0.0));
  }
  return tot;
}

// dot(W[j, :], h): the row is contiguous; four partial sums over k mod 4 and a tail into the first one (the summation
// order of every float64 MLP routine of this library, so results are bit-identical across them)
__device__ __forceinline__ double dot_row(const double *__restrict__ w, const double *__restrict__ h, int Kin) {
  double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
  int k = 0;
  for (; k + 8 <= Kin; k += 8) {
    p0 = fma(w[k + 0], h[k + 0], p0);
    p1 = fma(w[k + 1], h[k + 1], p1);
    p2 = fma(w[k + 2], h[k + 2], p2);
    p3 = fma(w[k + 3], h[k + 3], p3);
    p0 = fma(w[k + 4], h[k + 4], p0);
    p1 = fma(w[k + 5], h[k + 5], p1);
    p2 = fma(w[k + 6], h[k + 6], p2);
    p3 = fma(w[k + 7], h[k + 7], p3);
  }
  if (k + 4 <= Kin) {
    p0 = fma(w[k + 0], h[k + 0], p0);
    p1 = fma(w[k + 1], h[k + 1], p1);
    p2 = fma(w[k + 2], h[k + 2], p2);
    p3 = fma(w[k + 3], h[k + 3], p3);
    k += 4;
  }
  for (; k < Kin; ++k) p0 = fma(w[k], h[k], p0);
  return (p0 + p1) + (p2 + p3);
}

// barrier over the `nthr` threads of one line-search group (named barrier `id` >= 1), or a warp barrier
__device__ __forceinline__ void group_sync(int id, int nthr) {
  if (nthr == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthr) : "memory");
}

// One MLP forward for ONE sample by a group of `gthr` threads (`gl` = rank inside the group): h0 holds the z-scored
// input, h0/h1 ping-pong (shared memory), returns the buffer with the raw outputs  (mlp.py:55-59).
__device__ __forceinline__ const double *group_forward(const IlqrNet &net, const double *nb, double *h0, double *h1, int gl,
                                                       int gthr, int bar) {
  double *hin = h0, *hout = h1;
#pragma unroll 1
  for (int l = 0; l < net.n_layers; ++l) {
    const int Kin = net.dims[l], N = net.dims[l + 1], ws = net.wstride[l];
    const bool last = (l == net.n_layers - 1);
    const double *W = nb + net.woff[l], *B = nb + net.boff[l];
    for (int j = gl; j < N; j += gthr) {
      const double y = B[j] + dot_row(W + (size_t)j * ws, hin, Kin);
      hout[j] = last ? y : ampc_act<double>(net.act, y);
    }
    group_sync(bar, gthr);
    double *t2 = hin; hin = hout; hout = t2;
  }
  return hin;
}

// Jacobians of x' = x + dy(x,u) at (xs[i], us[i]) for i < H  ->  Jacs (H, nx, nx+nu)   (mlp.py:281-305)
// One WARP per sample (samples warp, warp + JAC_WARPS, ...), warp barriers only.  Forward-mode propagation of the
// [width x nin] panel through the layer stack.  Per layer a lane owns the output rows j = lane and lane + 32 (then
// lane + 64, ...) TOGETHER and keeps a CB-column strip of both in registers while it walks k: its weights are two
// contiguous rows, row k of the previous panel is a broadcast 16-byte load shared by both rows.
// `wk` = this warp's scratch: h0, h1, g (mw each, 16-byte aligned) and two panels (mw x jac_cols(nin) each).
constexpr int JAC_CB = 6;
__device__ __forceinline__ void jac_batch(const IlqrParams &P, const double *nb, const double *xs, const double *us, double *Jacs, double *wk,
                          int warp, int lane) {
  const IlqrNet &net = P.net;
  const int nx = P.nx, nu = P.nu, nin = nx + nu, H = P.H, mw = net.max_width, np = jac_cols(nin);
  const int mwp = (mw + 1) & ~1;                        // keeps the panels 16-byte aligned
  double *h0 = wk, *h1 = h0 + mwp, *g = h1 + mwp, *J0 = g + 2 * mwp, *J1 = J0 + (size_t)mw * np;
  const double *xu_mean = nb + net.xu_mean, *xu_std = nb + net.xu_std, *dy_std = nb + net.dy_std;
  if (warp < JAC_WARPS) {
    for (int s = warp; s < H; s += JAC_WARPS) {
      for (int j = lane; j < nin; j += 32) {
        const double v = j < nx ? xs[(size_t)s * nx + j] : us[(size_t)s * nu + (j - nx)];
        h0[j] = (v - xu_mean[j]) / xu_std[j];
      }
      __syncwarp();
      double *hin = h0, *hout = h1, *Jp = J0, *Jn = J1;
#pragma unroll 1
      for (int l = 0; l < net.n_layers; ++l) {
        const int Kin = net.dims[l], N = net.dims[l + 1], ws = net.wstride[l];
        const bool last = (l == net.n_layers - 1);
        const double *W = nb + net.woff[l], *B = nb + net.boff[l];
        for (int j = lane; j < N; j += 32) {
          const double y = B[j] + dot_row(W + (size_t)j * ws, hin, Kin);
          hout[j] = last ? y : ampc_act<double>(net.act, y);
          g[j] = last ? 1.0 : ampc_act_grad<double>(net.act, y);
        }
        __syncwarp();
        if (l == 0) {
          const double *W0s = nb + net.w0s;
          for (int e = lane; e < N * nin; e += 32) {
            const int j = e / nin, c = e - j * nin;
            Jn[(size_t)j * np + c] = W0s[e] * g[j];
          }
          if (np != nin) for (int j = lane; j < N; j += 32) Jn[(size_t)j * np + nin] = 0.0;
        } else if (N * np <= 256) {
          // narrow layer (the output layer): one (row, column) element per lane and pass instead of a whole row pair
          for (int e = lane; e < N * np; e += 32) {
            const int j = e / np, c = e - j * np;
            const double *wr = W + (size_t)j * ws, *r0 = Jp + c;
            double a0 = 0.0, a1 = 0.0;                   // even / odd k
            int k = 0;
            for (; k + 2 <= Kin; k += 2) { a0 = fma(wr[k], r0[(size_t)k * np], a0); a1 = fma(wr[k + 1], r0[(size_t)(k + 1) * np], a1); }
            if (k < Kin) a0 = fma(wr[k], r0[(size_t)k * np], a0);
            Jn[e] = (a0 + a1) * g[j];
          }
        } else {
          for (int j0 = lane; j0 < N; j0 += 64) {
            const int j1 = j0 + 32;
            const bool two = j1 < N;
            const double *wr0 = W + (size_t)j0 * ws, *wr1 = W + (size_t)(two ? j1 : j0) * ws;
            const double g0 = g[j0], g1 = two ? g[j1] : 0.0;
            for (int c0 = 0; c0 < np; c0 += JAC_CB) {
              // strips are JAC_CB wide; the last one may hang over the panel's row by up to JAC_CB - 2 columns: those
              // columns read the next row's leading entries (inside the panel except for its very last row, which has
              // mw - N rows of slack or the neighbouring buffer behind it) and are never stored
              double a00[JAC_CB], a01[JAC_CB], a10[JAC_CB], a11[JAC_CB];   // [row][k parity]
#pragma unroll
              for (int q = 0; q < JAC_CB; ++q) { a00[q] = 0.0; a01[q] = 0.0; a10[q] = 0.0; a11[q] = 0.0; }
              const double *r0 = Jp + c0;
              int k = 0;
              for (; k + 2 <= Kin; k += 2, r0 += 2 * np) {
                const double w00 = wr0[k], w01 = wr0[k + 1], w10 = wr1[k], w11 = wr1[k + 1];
#pragma unroll
                for (int q = 0; q < JAC_CB; q += 2) {
                  const double2 e = *reinterpret_cast<const double2 *>(r0 + q), o = *reinterpret_cast<const double2 *>(r0 + np + q);
                  a00[q] = fma(w00, e.x, a00[q]); a00[q + 1] = fma(w00, e.y, a00[q + 1]);
                  a01[q] = fma(w01, o.x, a01[q]); a01[q + 1] = fma(w01, o.y, a01[q + 1]);
                  a10[q] = fma(w10, e.x, a10[q]); a10[q + 1] = fma(w10, e.y, a10[q + 1]);
                  a11[q] = fma(w11, o.x, a11[q]); a11[q + 1] = fma(w11, o.y, a11[q + 1]);
                }
              }
              if (k < Kin) {
                const double w00 = wr0[k], w10 = wr1[k];
#pragma unroll
                for (int q = 0; q < JAC_CB; ++q) { a00[q] = fma(w00, r0[q], a00[q]); a10[q] = fma(w10, r0[q], a10[q]); }
              }
#pragma unroll
              for (int q = 0; q < JAC_CB; ++q)
                if (c0 + q < np) {
                  Jn[(size_t)j0 * np + c0 + q] = (a00[q] + a01[q]) * g0;
                  if (two) Jn[(size_t)j1 * np + c0 + q] = (a10[q] + a11[q]) * g1;
                }
            }
          }
        }
        __syncwarp();
        double *t2 = hin; hin = hout; hout = t2;
        double *t3 = Jp; Jp = Jn; Jn = t3;
      }
      double *dst = Jacs + (size_t)s * nx * nin;
      for (int e = lane; e < nx * nin; e += 32) {
        const int a = e / nin, c = e - a * nin;
        dst[e] = Jp[(size_t)a * np + c] * dy_std[a] + ((c == a) ? 1.0 : 0.0);
      }
      __syncwarp();
    }
  }
  __syncthreads();
}

// Line-search rollouts (ilqr.py:190-205) for all LS step sizes at once, like the reference's batched pred_batch, as a
// register-tiled product  out[alpha][j] = sum_k h[alpha][k] W[j][k].
//   main loop: 8 * AG warps = (alpha group ag) x (row half rh) x (K-quarter kq in 0..3); lane = output row j = 32 rh +
//     lane (+ 64, ...).  A thread accumulates, for its LA alphas, the partial sum over k == kq (mod 4) of its row --
//     exactly the four partial sums of dot_row.  Per k it loads its weight (conflict free with the odd row stride) and
//     its alphas' activations (stored [k][alpha]: the address is uniform across the warp) for LA multiply-adds.  With
//     AG = 1 (LS <= 10) every weight crosses the shared-memory pipe ONCE per layer and step;
//   combine / controls / state update: all LS_WARPS warps, one (alpha, output) element per thread, (alpha, j) per layer
//     computed once per call; the four K-quarters meet in shared memory and are combined in dot_row's order
//     (p0 + p1) + (p2 + p3).
// What the in-kernel phase counters said on the way (cycles per horizon step, 10 alphas, 5-64-64-4 network): one group
// of warps per alpha 8.8 k -- the whole network through the 128 B/clk shared-memory pipe ten times per step; a
// lane-level K split 12.4 k -- 16-byte activation loads with four addresses per warp; 16 main-loop warps in 4 alpha
// groups 10 k (main loops 5.8 k: the weights still cross the pipe four times; combine 2.5 k: a division per element).
template <int LA, int AG>
__device__ __forceinline__ void ls_rollouts(const IlqrParams &P, const double *nb, double *hT /* 2 x mw x LSP */,
                            double *part /* 4 x LSP x mw */, const double *states, const double *ctrls, const double *Ks,
                            const double *ks, double *ls_states, double *ls_ctrls, const double *c_alphas,
                            const double *c_umin, const double *c_umax, int tid) {
  const IlqrNet &net = P.net;
  const int nx = P.nx, nu = P.nu, n = nx + nu, H = P.H, LS = P.ls_max_iter, mw = net.max_width;
  constexpr int NTH = LS_WARPS * 32, LAP = (LA + 1) & ~1, LSP = AG * LAP;
  static_assert(8 * AG <= LS_WARPS, "main-loop warps");
  const int lane = tid & 31, wp = tid >> 5, kq = wp & 3, rh = (wp >> 2) & 1, ag = wp >> 3;
  const bool main_warp = wp < 8 * AG;
  const double *xu_mean = nb + net.xu_mean, *xu_std = nb + net.xu_std, *dy_mean = nb + net.dy_mean, *dy_std = nb + net.dy_std;
  double *hA = hT, *hB = hT + (size_t)mw * LSP;
  auto sync_ls = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NTH) : "memory"); };
  auto col_of = [&](int alpha) { const int g = alpha / LA; return g * LAP + (alpha - g * LA); };
  long long tm = clock64();
  unsigned long long lc[6] = {0, 0, 0, 0, 0, 0};         // thread 0: controls, main loops, barrier after main, combine, barrier after combine, update
  auto lapl = [&](int q) { const long long now = clock64(); lc[q] += (unsigned long long)(now - tm); tm = now; };
  // the (alpha, column) this thread prepares in the controls phase and the (alpha, state) it integrates
  const int cj = tid < LS * n ? tid / n : -1, cc = tid - (tid / n) * n;
  const int uj = tid < LS * nx ? tid / nx : -1, ua = tid - (tid / nx) * nx;
  for (int t = tid; t < LS * nx; t += NTH) {
    const int j = t / nx, a = t - j * nx;
    ls_states[(size_t)j * (H + 1) * nx + a] = P.x0[a];
  }
  for (int t = tid; t < 2 * mw * LSP; t += NTH) hT[t] = 0.0;       // padding alphas stay zero
  sync_ls();
#pragma unroll 1
  for (int i = 0; i < H; ++i) {
    // controls (ilqr.py:201-204) and the z-scored input, one thread per (alpha, input column)
    auto prep = [&](int j, int c) {
      const double *xi = ls_states + ((size_t)j * (H + 1) + i) * nx;
      double v;
      if (c < nx) {
        v = xi[c];
      } else {
        const int a = c - nx;
        const double *kr = Ks + ((size_t)i * nu + a) * nx, *x_ref = states + (size_t)i * nx;
        double fb = 0.0;
        for (int b = 0; b < nx; ++b) fb += kr[b] * (xi[b] - x_ref[b]);
        double u = c_alphas[j] * ks[(size_t)i * nu + a] + ctrls[(size_t)i * nu + a] + fb;
        if (P.bounded) u = fmin(fmax(u, c_umin[a]), c_umax[a]);     // np.clip, ilqr.py:203-204
        ls_ctrls[((size_t)j * H + i) * nu + a] = u;
        v = u;
      }
      hA[(size_t)c * LSP + col_of(j)] = (v - xu_mean[c]) / xu_std[c];
    };
    if (cj >= 0) prep(cj, cc);
    for (int t = tid + NTH; t < LS * n; t += NTH) prep(t / n, t % n);
    sync_ls();
    lapl(0);
    double *hin = hA, *hout = hB;
#pragma unroll 1
    for (int l = 0; l < net.n_layers; ++l) {
      const int Kin = net.dims[l], N = net.dims[l + 1], ws = net.wstride[l], Kin4 = Kin & ~3;
      const bool last = (l == net.n_layers - 1);
      const double *W = nb + net.woff[l], *B = nb + net.boff[l];
      if (main_warp) {
#pragma unroll 1
        for (int j0 = rh * 32; j0 < N; j0 += 64) {
          const int r0 = j0 + lane;
          const bool live = r0 < N;
          const double *wp0 = W + (size_t)(live ? r0 : 0) * ws;
          const double *hp = hin + ag * LAP;
          double p0[LA];
#pragma unroll
          for (int q = 0; q < LA; ++q) p0[q] = 0.0;
          auto step = [&](int k) {
            const double w = wp0[k];
            const double2 *hv = reinterpret_cast<const double2 *>(hp + (size_t)k * LSP);
#pragma unroll
            for (int q = 0; q < LAP / 2; ++q) {
              const double2 v = hv[q];
              p0[2 * q] = fma(w, v.x, p0[2 * q]);
              if (2 * q + 1 < LA) p0[2 * q + 1] = fma(w, v.y, p0[2 * q + 1]);
            }
          };
#pragma unroll 2
          for (int k = kq; k < Kin4; k += 4) step(k);
          if (kq == 0) for (int k = Kin4; k < Kin; ++k) step(k);     // dot_row's tail goes into the first partial sum
          if (live) {
            double *pp = part + ((size_t)kq * LSP + ag * LAP) * mw + r0;
#pragma unroll
            for (int q = 0; q < LA; ++q) pp[(size_t)q * mw] = p0[q];
          }
        }
      }
      lapl(1);
      sync_ls();
      lapl(2);
      {                                                             // (p0 + p1) + (p2 + p3), bias, activation
        const size_t qs = (size_t)LSP * mw;
        for (int jb = 0; jb < N; jb += 64) {
          const int j = jb + (tid & 63);
          if (j < N) {
            const double bj = B[j];
            for (int a = tid >> 6; a < LS; a += NTH / 64) {
              const int col = col_of(a);
              const double *q0 = part + (size_t)col * mw + j;
              const double y = bj + ((q0[0] + q0[qs]) + (q0[2 * qs] + q0[3 * qs]));
              hout[(size_t)j * LSP + col] = last ? y : ampc_act<double>(net.act, y);
            }
          }
        }
      }
      lapl(3);
      sync_ls();
      lapl(4);
      double *t2 = hin; hin = hout; hout = t2;
    }
    if (uj >= 0) {                                                   // mlp.py:235-236
      const double *xi = ls_states + ((size_t)uj * (H + 1) + i) * nx;
      ls_states[((size_t)uj * (H + 1) + i + 1) * nx + ua] = xi[ua] + (hin[(size_t)ua * LSP + col_of(uj)] * dy_std[ua] + dy_mean[ua]);
    }
    for (int t = tid + NTH; t < LS * nx; t += NTH) {
      const int j = t / nx, a = t - j * nx;
      const double *xi = ls_states + ((size_t)j * (H + 1) + i) * nx;
      ls_states[((size_t)j * (H + 1) + i + 1) * nx + a] = xi[a] + (hin[(size_t)a * LSP + col_of(j)] * dy_std[a] + dy_mean[a]);
    }
    sync_ls();
    lapl(5);
  }
  if (tid == 0) for (int q = 0; q < 6; ++q) P.prof[8 + q] += lc[q];
}

// D (8x8) += A (8x4, row-major) . B (4x8, column-major), float64, one warp.  Fragments (lane = 4 g + t): a = A[g][t],
// b = B[t][g], c0 / c1 = C[g][2 t] / C[g][2 t + 1].
__device__ __forceinline__ void dmma_884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Activations other than ReLU out of line: the horizon loop of ls_rollouts_mma is latency bound, and the inlined
// float64 tanh / exp / expm1 bodies (twice per call site) were most of its instruction footprint.
__device__ __noinline__ double ilqr_act_other(int act, double y) { return ampc_act<double>(act, y); }
__device__ __forceinline__ double ilqr_act(int act, double y) {
  if (act == AMPC_ACT_RELU) return y > 0.0 ? y : 0.0;
  return ilqr_act_other(act, y);
}

// 32-bit shared-memory addressing for ls_rollouts_mma: every array it touches is shared-memory resident (the caller only
// takes this path in resident mode), and the horizon loop is a latency-bound chain -- with generic double pointers the
// compiler rebuilt 64-bit addresses (and the shared window base from SR_CgaCtaId) in front of almost every load.
__device__ __forceinline__ uint32_t sh32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}

// Everything ls_rollouts_mma needs, in shared memory (filled once per solve by the kernel): the routine is compiled OUT
// OF LINE so that it gets a register allocation of its own -- inlined into the solve kernel it competed with the other
// phases for the 128 registers, the compiler re-derived its loop-invariant addresses from special registers and the
// constant bank in front of the loads of every layer, and the other phases spilled more.
struct LsArgs {
  int nx, nu, H, S, mwp, L, act, bounded, profile;
  int KS[MAXL], MT[MAXL], N[MAXL], Kin[MAXL], ws[MAXL];      // per layer: k-steps, row tiles, outputs, inputs, row stride of W
  uint32_t a_W[MAXL], a_B[MAXL], a_Wf[MAXL];                 // shared addresses: row-major weights, biases, fragment image
  uint32_t a_xu_mean, a_xu_std, a_dy_mean, a_dy_std;
  uint32_t a_scr, a_states, a_ctrls, a_Ks, a_ks, a_ls_states, a_ls_ctrls, a_alphas, a_umin, a_umax;
  const double *x0;                                          // global
  unsigned long long *prof;                                  // global
  // jac_batch_mma
  uint32_t a_W0s, a_G, a_PA, a_PB, a_tab, a_Jacs;            // W0 / xu_std (row-major), act' per hidden layer, panels, column -> step table
  int SP, ncols;                                             // panel row stride (doubles), 8 * (nx + nu) panel columns per chunk
  int wf_persistent;                                         // the weight image is built once per solve (jac_batch_mma in use)
};

// Line-search rollouts on the FP64 tensor-core path: same contract as ls_rollouts for the `na` (<= 8) step sizes
// a0 .. a0 + na - 1, which occupy the 8 columns of one alpha tile.  (The caller evaluates the step sizes eight at a time
// and stops as soon as the acceptance rule of ilqr.py:208-225, which walks them in order, has made its choice: the
// outputs are those of the reference's all-at-once batch, the rollouts nobody reads are not computed.)
// Per horizon step and layer the product  out[j][alpha] = sum_k W[j][k] h[alpha][k]  is tiled as
// (8 output rows) x (8 alphas) x (4 k)  mma.m8n8k4: warp w of the 8 main warps owns the row tiles w, w + 8, ...; the
// weights come from a FRAGMENT-ORDERED image built once per call in the phase scratch (tile, k-step, lane -> one
// double: a conflict-free 256-byte load per warp), the activations sit as [k][S] with S = 4 (mod 16) doubles so that the
// (4 k) x (8 alphas) operand of a warp is conflict free as well.  That removes what bounded ls_rollouts (one
// shared-memory operand per multiply-add per lane: the register write-back of warp-wide loads) -- here one operand pair
// feeds 256 multiply-adds -- and the K-quarter partials with their combine pass: four accumulators per tile take the
// k-steps 0..3 (mod 4) and are summed as (p0 + p1) + (p2 + p3).  The input z-score's division is folded into the image
// of the first layer (columns scaled by 1 / xu_std), the mean is subtracted when the input is staged.  An output layer of
// one or two row tiles (nx <= 16) is split over K instead: warp 4 tile + q takes the k-steps q (mod 4), i.e. one of the
// four partial sums, and the state update combines them.  Padding alphas / rows / k carry finite garbage times zero.
// Every array is shared-memory resident (the caller takes this path in resident mode only) and addressed with 32-bit
// shared addresses; per-thread roles (which (alpha, column) a thread stages, which (alpha, state) it integrates), their
// addresses and the per-layer constants are fixed before the horizon loop, inside it addresses advance by constants.
// NXT = nx at compile time (0: run time).
// (scripts/dmma_microbench.cu, profiles/r02b_dmma_microbench.txt: 26 cycles latency, one mma per 16 cycles and SM
// sub-partition = 64 multiply-adds per clock and SM, the DFMA rate -- the gain is in the operand traffic.)
// The routine is two loops over the horizon that meet at the barriers, each compiled on its own (a combined body kept
// ~120 values alive and spilled; with 227 KB of shared memory the L1 that backs local memory is a few KB, so a spill
// costs an L2 round trip):
//   ls_mma_io      warps 8..15: stage the inputs of step i (controls, centred state), integrate the state after it;
//   ls_mma_layers  warps 0..7 : the layers of step i on mma fragments.
// Barrier 1 (all 16 warps) separates staging | layers | update; barrier 2 (8 warps) the layers among themselves.
template <int NXT>
__device__ __noinline__ void ls_mma_io(const LsArgs *A, int a0, int na) {
  constexpr int NTH = LS_WARPS * 32, NMAIN = 8;
  const int tid = (int)threadIdx.x - NMAIN * 32;         // 0 .. 255 among the io warps
  const int nx = NXT ? NXT : A->nx, nu = A->nu, n = nx + nu, H = A->H, S = A->S, L = A->L;
  const uint32_t S8 = (uint32_t)S * 8u;
  const uint32_t a_hA = A->a_scr, a_hB = a_hA + (uint32_t)A->mwp * S8, a_part = a_hB + (uint32_t)A->mwp * S8;
  auto sync_ls = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NTH) : "memory"); };
  // roles and addresses (the host takes this path only when 8 (nx + nu) <= 256)
  const bool stager = tid < na * n;
  const int pj = stager ? tid / n : 0, pc = stager ? tid - pj * n : 0;
  const bool is_ctl = stager && pc >= nx;
  const int pa = is_ctl ? pc - nx : 0;
  const double p_mean = lds64(A->a_xu_mean + (uint32_t)pc * 8u);
  const uint32_t p_dst = a_hA + (uint32_t)(pc * S + pj) * 8u;
  uint32_t p_row = A->a_ls_states + (uint32_t)((a0 + pj) * (H + 1) * nx) * 8u;       // this alpha's state at step i
  uint32_t p_ref = A->a_states, p_kr = A->a_Ks + (uint32_t)(pa * nx) * 8u, p_ks = A->a_ks + (uint32_t)pa * 8u,
           p_ct = A->a_ctrls + (uint32_t)pa * 8u, p_lsc = A->a_ls_ctrls + (uint32_t)((a0 + pj) * H * nu + pa) * 8u;
  const double p_alpha = lds64(A->a_alphas + (uint32_t)(a0 + pj) * 8u), p_umin = lds64(A->a_umin + (uint32_t)pa * 8u),
               p_umax = lds64(A->a_umax + (uint32_t)pa * 8u);
  const bool bounded = A->bounded != 0;
  const bool integrator = tid < na * nx;
  const int uj = integrator ? tid / nx : 0, ua = integrator ? tid - uj * nx : 0;
  const double u_std = lds64(A->a_dy_std + (uint32_t)ua * 8u), u_mean = lds64(A->a_dy_mean + (uint32_t)ua * 8u),
               u_bias = lds64(A->a_B[L - 1] + (uint32_t)ua * 8u);
  uint32_t u_x = A->a_ls_states + (uint32_t)((a0 + uj) * (H + 1) * nx + ua) * 8u;
  const bool ksplit = A->MT[L - 1] * 4 <= NMAIN;
  const uint32_t u_part = a_part + (uint32_t)(((ua >> 3) * 4) * 64 + (ua & 7) * 8 + uj) * 8u;
  const uint32_t u_y = (((L & 1) ? a_hB : a_hA)) + (uint32_t)(ua * S + uj) * 8u;    // where the last layer's output lands
  const uint32_t nx8 = (uint32_t)nx * 8u, nu8 = (uint32_t)nu * 8u, knx8 = (uint32_t)(nu * nx) * 8u;
  if (integrator) sts64(u_x, A->x0[ua]);                 // every alpha starts from x0
  sync_ls();                                             // (pairs with the entry barrier of ls_mma_layers)
#pragma unroll 1
  for (int i = 0; i < H; ++i) {
    // controls (ilqr.py:201-204) and the centred input, one thread per (alpha, input column)
    if (stager) {
      double v;
      if (!is_ctl) {
        v = lds64(p_row + (uint32_t)pc * 8u);
      } else {
        double fb = 0.0;
        if constexpr (NXT > 0) {
          double kr[NXT], xi[NXT], xr[NXT];
#pragma unroll
          for (int b = 0; b < NXT; ++b) { kr[b] = lds64(p_kr + b * 8u); xi[b] = lds64(p_row + b * 8u); xr[b] = lds64(p_ref + b * 8u); }
#pragma unroll
          for (int b = 0; b < NXT; ++b) fb = fma(kr[b], xi[b] - xr[b], fb);
        } else {
          for (int b = 0; b < nx; ++b) fb = fma(lds64(p_kr + b * 8u), lds64(p_row + b * 8u) - lds64(p_ref + b * 8u), fb);
        }
        double u = fma(p_alpha, lds64(p_ks), lds64(p_ct)) + fb;
        if (bounded) u = fmin(fmax(u, p_umin), p_umax);               // np.clip, ilqr.py:203-204
        sts64(p_lsc, u);
        v = u;
        p_kr += knx8; p_ks += nu8; p_ct += nu8; p_lsc += nu8;
      }
      sts64(p_dst, v - p_mean);
      p_row += nx8; p_ref += nx8;
    }
    sync_ls();                                           // inputs staged -> layers
    sync_ls();                                           // layers done
    // mlp.py:235-236; y = b + (p0 + p1) + (p2 + p3) when the output layer was split over K
    if (integrator) {
      double y;
      if (ksplit) y = u_bias + ((lds64(u_part) + lds64(u_part + 512u)) + (lds64(u_part + 1024u) + lds64(u_part + 1536u)));
      else y = lds64(u_y);
      sts64(u_x + nx8, lds64(u_x) + fma(y, u_std, u_mean));
      u_x += nx8;
    }
    sync_ls();                                           // state integrated
  }
}

__device__ __noinline__ void ls_mma_layers(const LsArgs *A) {
  constexpr int NTH = LS_WARPS * 32, NMAIN = 8;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  const int H = A->H, S = A->S, L = A->L;
  const uint32_t S8 = (uint32_t)S * 8u;
  const uint32_t a_hA = A->a_scr, a_hB = a_hA + (uint32_t)A->mwp * S8, a_part = a_hB + (uint32_t)A->mwp * S8;
  auto sync_ls = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NTH) : "memory"); };
  auto sync_main = [&]() { asm volatile("bar.sync 2, %0;" ::"n"(NMAIN * 32) : "memory"); };
  const bool prof = A->profile != 0;
  long long tm = prof ? clock64() : 0;
  unsigned long long lc[6] = {0, 0, 0, 0, 0, 0};         // thread 0: staging (io warps), layer 0, output layer, other layers, -, update (io warps)
  auto lapl = [&](int q) { if (prof) { const long long now = clock64(); lc[q] += (unsigned long long)(now - tm); tm = now; } };
  const bool ksplit = A->MT[L - 1] * 4 <= NMAIN;         // output layer split over K
  const int act = A->act;
  const uint32_t hb_off = (uint32_t)(t4 * S + g) * 8u, out_off = (uint32_t)(g * S + 2 * t4) * 8u, lane8 = (uint32_t)lane * 8u;
  sync_ls();
#pragma unroll 1
  for (int i = 0; i < H; ++i) {
    sync_ls();                                           // inputs staged
    lapl(0);
    uint32_t hin = a_hA, hout = a_hB;
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const int KS = A->KS[l], MT = A->MT[l];
      const bool last = (l == L - 1);
      const uint32_t hb = hin + hb_off, wl = A->a_Wf[l] + lane8;
      if (last && ksplit) {
        // output layer, K split: this warp = (row tile, partial q); a chain of KS / 4 dependent mma
        const int mt = wp >> 2, q = wp & 3;
        if (mt < MT) {
          double c0 = 0.0, c1 = 0.0;
          const uint32_t wa = wl + (uint32_t)(mt * KS) * 256u;
#pragma unroll 4
          for (int ks = q; ks < KS; ks += 4) dmma_884(c0, c1, lds64(wa + (uint32_t)ks * 256u), lds64(hb + (uint32_t)ks * 4u * S8));
          sts128(a_part + (uint32_t)((mt * 4 + q) * 64 + g * 8 + 2 * t4) * 8u, c0, c1);
        }
        lapl(2);
        break;                                                      // the state update combines the partials behind the barrier
      }
      const int N = A->N[l];
      const uint32_t a_B = A->a_B[l];
#pragma unroll 1
      for (int mt = wp; mt < MT; mt += NMAIN) {
        double acc[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = 0.0;
        const uint32_t wa = wl + (uint32_t)(mt * KS) * 256u;
        // software pipeline: the operands of the next four k-steps are in flight while this group's mma issue
        double a_cur[4], b_cur[4], a_nxt[4], b_nxt[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { a_cur[q] = lds64(wa + q * 256u); b_cur[q] = lds64(hb + (uint32_t)q * 4u * S8); }
#pragma unroll 1
        for (int ks = 0; ks < KS; ks += 4) {
          const bool more = ks + 4 < KS;
          if (more) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              a_nxt[q] = lds64(wa + (uint32_t)(ks + 4 + q) * 256u);
              b_nxt[q] = lds64(hb + (uint32_t)(ks + 4 + q) * 4u * S8);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) dmma_884(acc[q][0], acc[q][1], a_cur[q], b_cur[q]);
          if (more) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { a_cur[q] = a_nxt[q]; b_cur[q] = b_nxt[q]; }
          }
        }
        const int j = 8 * mt + g;
        const double bj = j < N ? lds64(a_B + (uint32_t)j * 8u) : 0.0;
        double y0 = bj + ((acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]));
        double y1 = bj + ((acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]));
        if (!last) { y0 = ilqr_act(act, y0); y1 = ilqr_act(act, y1); }
        sts128(hout + (uint32_t)(8 * mt) * S8 + out_off, y0, y1);
      }
      sync_main();
      if (l == 0) lapl(1); else lapl(last ? 2 : 3);
      const uint32_t t2 = hin; hin = hout; hout = t2;
    }
    sync_ls();                                           // layers done -> update
    sync_ls();                                           // state integrated
    lapl(5);
  }
  if (prof && tid == 0) for (int q = 0; q < 6; ++q) A->prof[8 + q] += lc[q];
}

// Caller side: all LS_WARPS warps.  Builds the fragment-ordered weight image (the per-warp Jacobian panels or the tensor-
// core Jacobian refresh have used the phase scratch since the last call) and hands the two halves of the CTA their loops.
// Fragment-ordered weight image of every layer: Wf_l[(tile * KS + ks) * 32 + lane] = W_l[8 tile + lane / 4][4 ks + lane % 4],
// zero padded; layer 0 carries 1 / xu_std.  By the first NTH threads of the CTA.
__device__ __forceinline__ void ls_build_wf(const LsArgs *A, int NTH) {
  const int tid = threadIdx.x;
  if (tid >= NTH) return;
  for (int l = 0; l < A->L; ++l) {
    const int Kin = A->Kin[l], N = A->N[l], ws = A->ws[l], KS = A->KS[l], cnt = A->MT[l] * KS * 32;
    const uint32_t W = A->a_W[l], dst = A->a_Wf[l];
    for (int e = tid; e < cnt; e += NTH) {
      const int ln = e & 31, ks = (e >> 5) % KS, mt = (e >> 5) / KS, j = 8 * mt + (ln >> 2), k = 4 * ks + (ln & 3);
      double w = (j < N && k < Kin) ? lds64(W + (uint32_t)(j * ws + k) * 8u) : 0.0;
      if (l == 0 && k < Kin) w /= lds64(A->a_xu_std + (uint32_t)k * 8u);   // z = (v - mean) / std  ->  (v - mean) . (W / std)
      sts64(dst + (uint32_t)e * 8u, w);
    }
  }
}

template <int NXT>
__device__ __forceinline__ void ls_rollouts_mma(const LsArgs *A, int a0, int na) {
  constexpr int NTH = LS_WARPS * 32;
  const int tid = threadIdx.x;
  if (!A->wf_persistent) {                               // the per-warp Jacobian panels have used the phase scratch since the last call
    ls_build_wf(A, NTH);
    for (int t = tid; t < 2 * A->mwp * A->S; t += NTH) sts64(A->a_scr + (uint32_t)t * 8u, 0.0);
  }
  if (tid < 8 * 32) ls_mma_layers(A);
  else ls_mma_io<NXT>(A, a0, na);
}

__device__ __noinline__ double ilqr_act_grad_other(int act, double y) { return ampc_act_grad<double>(act, y); }
__device__ __forceinline__ double ilqr_act_grad(int act, double y) {
  if (act == AMPC_ACT_RELU) return y > 0.0 ? 1.0 : 0.0;
  return ilqr_act_grad_other(act, y);
}

// Jacobian refresh on the FP64 tensor-core path (same contract as jac_batch, resident mode only): the horizon steps are
// independent here, so they are the COLUMNS of the products -- eight steps per chunk.  Per chunk:
//   forward pass of the hidden layers for the 8 steps (one 8-column tile, as in ls_rollouts_mma) -> activations and the
//     activation derivatives g_l[j][step];
//   first panel  P_0[j][(step, c)] = W_0[j][c] / xu_std[c] * g_0[j][step]  (element-wise, mlp.py:298);
//   P_l = g_l . (W_l P_{l-1})  for the layers above: (8 rows) x (8 columns) x (4 k) mma tiles over the 8 (nx + nu) panel
//     columns, tile pairs dealt to the 16 warps; the weights are the line search's fragment-ordered image;
//   Jacs[step][a][c] = P_last[a][(step, c)] dy_std[a] + [a == c].
// FP64 floor of the refresh at the 5-64-64-4 network: ~18 k cycles (64 multiply-adds per clock); the one-warp-per-step
// routine takes ~107 k (operand loads, as in the old line search).
__device__ __noinline__ void jac_batch_mma(const LsArgs *A, uint32_t a_xs, uint32_t a_us) {
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  const int nx = A->nx, nu = A->nu, nin = nx + nu, H = A->H, S = A->S, L = A->L, SP = A->SP, ncols = A->ncols;
  constexpr int NTH = NT, NW = NT / 32;
  const uint32_t S8 = (uint32_t)S * 8u, SP8 = (uint32_t)SP * 8u;
  const uint32_t a_hA = A->a_scr, a_hB = a_hA + (uint32_t)A->mwp * S8;
  const int act = A->act;
  // (weight image, zeroed buffers and the column -> step table: set up once per solve by the kernel, they persist -- no
  // phase between two calls writes that part of the phase scratch when this routine is in use)
  const int NTP = (ncols + 7) >> 3;                      // column tiles of a panel
#pragma unroll 1
  for (int c0 = 0; c0 < H; c0 += 8) {
    const int ns = H - c0 < 8 ? H - c0 : 8;
    // centred inputs of the chunk's steps (columns beyond ns keep the previous chunk's finite values)
    for (int t = tid; t < ns * nin; t += NTH) {
      const int sI = t / nin, c = t - sI * nin;
      const double v = c < nx ? lds64(a_xs + (uint32_t)((c0 + sI) * nx + c) * 8u) : lds64(a_us + (uint32_t)((c0 + sI) * nu + (c - nx)) * 8u);
      sts64(a_hA + (uint32_t)(c * S + sI) * 8u, v - lds64(A->a_xu_mean + (uint32_t)c * 8u));
    }
    __syncthreads();
    // ---- forward pass of the hidden layers: activations and derivatives
    uint32_t hin = a_hA, hout = a_hB;
#pragma unroll 1
    for (int l = 0; l + 1 < L; ++l) {
      const int KS = A->KS[l], MT = A->MT[l], N = A->N[l];
      const uint32_t hb = hin + (uint32_t)(t4 * S + g) * 8u, a_G = A->a_G + (uint32_t)(l * A->mwp * 8) * 8u;
#pragma unroll 1
      for (int mt = wp; mt < MT; mt += NW) {
        double acc[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = 0.0;
        const uint32_t wa = A->a_Wf[l] + (uint32_t)(mt * KS * 32 + lane) * 8u;
#pragma unroll 1
        for (int ks = 0; ks < KS; ks += 4) {
          double a4[4], b4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) { a4[q] = lds64(wa + (uint32_t)(ks + q) * 256u); b4[q] = lds64(hb + (uint32_t)(ks + q) * 4u * S8); }
#pragma unroll
          for (int q = 0; q < 4; ++q) dmma_884(acc[q][0], acc[q][1], a4[q], b4[q]);
        }
        const int j = 8 * mt + g;
        const double bj = j < N ? lds64(A->a_B[l] + (uint32_t)j * 8u) : 0.0;
        const double y0 = bj + ((acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]));
        const double y1 = bj + ((acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]));
        sts128(hout + (uint32_t)(j * S + 2 * t4) * 8u, ilqr_act(act, y0), ilqr_act(act, y1));
        sts128(a_G + (uint32_t)(j * 8 + 2 * t4) * 8u, ilqr_act_grad(act, y0), ilqr_act_grad(act, y1));
      }
      __syncthreads();
      const uint32_t t2 = hin; hin = hout; hout = t2;
    }
    // ---- first panel (mlp.py:298)
    {
      const int N = A->N[0];
      for (int col = lane; col < ncols; col += 32) {
        int sI;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(sI) : "r"(A->a_tab + (uint32_t)col * 4u) : "memory");
        const int c = col - sI * nin;
        for (int j = wp; j < N; j += NW)
          sts64(A->a_PA + (uint32_t)(j * SP + col) * 8u,
                lds64(A->a_W0s + (uint32_t)(j * nin + c) * 8u) * lds64(A->a_G + (uint32_t)(j * 8 + sI) * 8u));
      }
    }
    __syncthreads();
    // ---- panels of the layers above
    uint32_t pin = A->a_PA, pout = A->a_PB;
#pragma unroll 1
    for (int l = 1; l < L; ++l) {
      const int KS = A->KS[l], MT = A->MT[l];
      const bool last = (l == L - 1);
      const uint32_t a_G = A->a_G + (uint32_t)(l * A->mwp * 8) * 8u;
#pragma unroll 1
      for (int pr = wp; pr < MT * NTP; pr += NW) {
        const int mt = pr % MT, nt = pr / MT;
        double acc[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = 0.0;
        const uint32_t wa = A->a_Wf[l] + (uint32_t)(mt * KS * 32 + lane) * 8u;
        const uint32_t pb = pin + (uint32_t)(t4 * SP + 8 * nt + g) * 8u;
        double a_cur[4], b_cur[4], a_nxt[4], b_nxt[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { a_cur[q] = lds64(wa + q * 256u); b_cur[q] = lds64(pb + (uint32_t)q * 4u * SP8); }
#pragma unroll 1
        for (int ks = 0; ks < KS; ks += 4) {
          const bool more = ks + 4 < KS;
          if (more) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              a_nxt[q] = lds64(wa + (uint32_t)(ks + 4 + q) * 256u);
              b_nxt[q] = lds64(pb + (uint32_t)(ks + 4 + q) * 4u * SP8);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) dmma_884(acc[q][0], acc[q][1], a_cur[q], b_cur[q]);
          if (more) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { a_cur[q] = a_nxt[q]; b_cur[q] = b_nxt[q]; }
          }
        }
        const int j = 8 * mt + g, col = 8 * nt + 2 * t4;
        double v0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
        double v1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
        if (!last) {
          int s0, s1;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(s0) : "r"(A->a_tab + (uint32_t)(col < ncols ? col : 0) * 4u) : "memory");
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(s1) : "r"(A->a_tab + (uint32_t)(col + 1 < ncols ? col + 1 : 0) * 4u) : "memory");
          v0 *= lds64(a_G + (uint32_t)(j * 8 + s0) * 8u);
          v1 *= lds64(a_G + (uint32_t)(j * 8 + s1) * 8u);
        }
        sts128(pout + (uint32_t)(j * SP + col) * 8u, v0, v1);
      }
      __syncthreads();
      const uint32_t t2 = pin; pin = pout; pout = t2;
    }
    // ---- mlp.py:300-305: scale by dy_std, add the identity of x' = x + dy
    for (int e = tid; e < ns * nx * nin; e += NTH) {
      const int sI = e / (nx * nin), r = e - sI * nx * nin, a = r / nin, c = r - a * nin;
      sts64(A->a_Jacs + (uint32_t)(((c0 + sI) * nx + a) * nin + c) * 8u,
            lds64(pin + (uint32_t)(a * SP + sI * nin + c) * 8u) * lds64(A->a_dy_std + (uint32_t)a * 8u) + ((c == a) ? 1.0 : 0.0));
    }
    __syncthreads();
  }
}

// Backward Riccati pass (ilqr.py:159-187) by ONE warp with warp barriers only.  A lone warp retires a dependent
// instruction every ~7-10 cycles, so the recursion costs what it executes: NX / NU > 0 fix the dimensions at compile time
// (loops unroll, index arithmetic folds; the cartpole shape 4 / 1 is instantiated), 0 = run-time dimensions.  Each
// lane's (row, column) per phase is fixed over the horizon and is looked up once (when a phase has more than 32 elements
// the lane strides over the rest).
struct BackwardSmem {
  double *Ct, *Fs, *V0, *V1, *v0, *v1, *T, *Qt, *qt, *K, *k, *LU;
  int *tab_n, *tab_x, *piv;
};
template <int NX, int NU>
__device__ __forceinline__ void backward_pass(const IlqrParams &P, const BackwardSmem &B, const double *states,
                                              const double *ctrls, const double *Jacs, double *Ks, double *ks,
                                              const double *c_goal, double *lin_out, double *quad_out, int lane) {
  const int nx = NX ? NX : P.nx, nu = NU ? NU : P.nu, n = nx + nu, H = P.H;
  double *const s_Ct = B.Ct, *const s_Fs = B.Fs, *const s_V0 = B.V0, *const s_V1 = B.V1, *const s_v0 = B.v0, *const s_v1 = B.v1,
         *const s_T = B.T, *const s_Qt = B.Qt, *const s_qt = B.qt, *const s_K = B.K, *const s_k = B.k, *const s_LU = B.LU;
  int *const s_tab_n = B.tab_n, *const s_tab_x = B.tab_x, *const s_piv = B.piv;
  double s_lin = 0.0, s_quad = 0.0;
  double *Vn = s_V0, *Vnn = s_V1, *vn = s_v0, *vnn = s_v1;
  for (int t = lane; t < nx * nx; t += 32) Vn[t] = s_Fs[t];
  for (int a = lane; a < nx; a += 32) {
    double acc = 0.0;
    for (int b = 0; b < nx; ++b) acc += s_Fs[a * nx + b] * states[(size_t)H * nx + b];   // no goal: cost.py:208
    vn[a] = acc;
  }
  double lin_acc = 0.0, quad_acc = 0.0;               // lane 0's running sums (ilqr.py:178-179)
  // element e of T (nx x n): T[a][c] = Vn[a,:] . J[:,c]
  auto do_T = [&](int e, int rc, const double *Vc, const double *J) {
    const int a = rc >> 16, c = rc & 0xffff;
    const double *vr = Vc + a * nx, *jc = J + c;
    double acc = 0.0;
    for (int b = 0; b < nx; ++b) acc += vr[b] * jc[b * n];
    s_T[e] = acc;
  };
  // element e of [Qt (n x n) | qt (n)]: Qt = Ct + J^T T ; qt = ct + J^T vn
  auto do_Q = [&](int e, int rc, const double *J, const double *vc, const double *xt, const double *ut) {
    if (e < n * n) {
      const int r = rc >> 16, c = rc & 0xffff;
      const double *jr = J + r, *tc = s_T + c;
      double acc = 0.0;
      for (int a = 0; a < nx; ++a) acc += jr[a * n] * tc[a * n];
      s_Qt[e] = s_Ct[e] + acc;
    } else {
      const int r = e - n * n;
      const double *cr = s_Ct + r * n;
      double ct = 0.0;
      if (r < nx) { for (int b = 0; b < nx; ++b) ct += cr[b] * (xt[b] - c_goal[b]); }
      else { for (int b = 0; b < nu; ++b) ct += cr[nx + b] * ut[b]; }
      double acc = 0.0;
      for (int a = 0; a < nx; ++a) acc += J[a * n + r] * vc[a];
      s_qt[r] = ct + acc;
    }
  };
  // element e of [V' (nx x nx) | v' (nx)]  (ilqr.py:186-187)
  auto do_V = [&](int e, int ab, double *Vo, double *vo) {
    if (e < nx * nx) {
      const int a = ab >> 16, b = ab & 0xffff;
      double acc = s_Qt[a * n + b];
      for (int r = 0; r < nu; ++r) acc += s_Qt[a * n + nx + r] * s_K[r * nx + b];
      for (int r = 0; r < nu; ++r) acc += s_K[r * nx + a] * s_Qt[(nx + r) * n + b];
      for (int r = 0; r < nu; ++r) {
        double row = 0.0;
        for (int c = 0; c < nu; ++c) row += s_Qt[(nx + r) * n + nx + c] * s_K[c * nx + b];
        acc += s_K[r * nx + a] * row;
      }
      Vo[e] = acc;
    } else {
      const int a = e - nx * nx;
      double acc = s_qt[a];
      for (int r = 0; r < nu; ++r) acc += s_Qt[a * n + nx + r] * s_k[r];
      for (int r = 0; r < nu; ++r) {
        double inner = s_qt[nx + r];
        for (int c = 0; c < nu; ++c) inner += s_Qt[(nx + r) * n + nx + c] * s_k[c];
        acc += s_K[r * nx + a] * inner;
      }
      vo[a] = acc;
    }
  };
  const int nT = nx * n, nQ = n * n + n, nV = nx * nx + nx;
  const int rcT = lane < nT ? s_tab_n[lane] : 0;
  const int rcQ = lane < n * n ? s_tab_n[lane] : 0;
  const int abV = lane < nx * nx ? s_tab_x[lane] : 0;
  __syncwarp();
  for (int t = H; t >= 1; --t) {
    const double *J = Jacs + (size_t)(t - 1) * nx * n;
    const double *xt = states + (size_t)(t - 1) * nx, *ut = ctrls + (size_t)(t - 1) * nu;
    if (lane < nT) do_T(lane, rcT, Vn, J);
    for (int e = lane + 32; e < nT; e += 32) do_T(e, s_tab_n[e], Vn, J);
    __syncwarp();
    if (lane < nQ) do_Q(lane, rcQ, J, vn, xt, ut);
    for (int e = lane + 32; e < nQ; e += 32) do_Q(e, e < n * n ? s_tab_n[e] : 0, J, vn, xt, ut);
    __syncwarp();
    if (nu == 1) {                                    // scalar Quu: K = -Qux / Quu, k = -qu / Quu (what gesv does for 1 x 1)
      const double quu = s_Qt[nx * n + nx];
      for (int col = lane; col <= nx; col += 32) {
        const double y = ((col < nx) ? s_Qt[nx * n + col] : s_qt[nx]) / quu;
        if (col < nx) s_K[col] = -y; else s_k[0] = -y;
      }
    } else {
      if (lane == 0) {                                // LU of Quu with partial pivoting (LAPACK gesv)
        for (int r = 0; r < nu; ++r) for (int c = 0; c < nu; ++c) s_LU[r * nu + c] = s_Qt[(nx + r) * n + nx + c];
        for (int c = 0; c < nu; ++c) {
          int pr = c; double best = fabs(s_LU[c * nu + c]);
          for (int r = c + 1; r < nu; ++r) if (fabs(s_LU[r * nu + c]) > best) { best = fabs(s_LU[r * nu + c]); pr = r; }
          s_piv[c] = pr;
          if (pr != c) for (int q = 0; q < nu; ++q) { double tmp = s_LU[c * nu + q]; s_LU[c * nu + q] = s_LU[pr * nu + q]; s_LU[pr * nu + q] = tmp; }
          for (int r = c + 1; r < nu; ++r) {
            s_LU[r * nu + c] /= s_LU[c * nu + c];
            for (int q = c + 1; q < nu; ++q) s_LU[r * nu + q] -= s_LU[r * nu + c] * s_LU[c * nu + q];
          }
        }
      }
      __syncwarp();
      for (int col = lane; col <= nx; col += 32) {    // K = -Quu^-1 Qux, k = -Quu^-1 qu: one right-hand side per lane
        double y[MAX_NU];
        for (int r = 0; r < nu; ++r) y[r] = (col < nx) ? s_Qt[(nx + r) * n + col] : s_qt[nx + r];
        for (int c = 0; c < nu; ++c) { const int pc = s_piv[c]; if (pc != c) { double tmp = y[c]; y[c] = y[pc]; y[pc] = tmp; } }
        for (int r = 1; r < nu; ++r) for (int q = 0; q < r; ++q) y[r] -= s_LU[r * nu + q] * y[q];
        for (int r = nu - 1; r >= 0; --r) { for (int q = r + 1; q < nu; ++q) y[r] -= s_LU[r * nu + q] * y[q]; y[r] /= s_LU[r * nu + r]; }
        for (int r = 0; r < nu; ++r) { if (col < nx) s_K[r * nx + col] = -y[r]; else s_k[r] = -y[r]; }
      }
    }
    __syncwarp();
    if (lane == 0) {
      double lin = 0.0, quad = 0.0;
      for (int r = 0; r < nu; ++r) {
        lin += s_qt[nx + r] * s_k[r];
        double row = 0.0;
        for (int c = 0; c < nu; ++c) row += s_Qt[(nx + r) * n + nx + c] * s_k[c];
        quad += s_k[r] * row;
      }
      lin_acc += lin; quad_acc += quad;
    }
    for (int e = lane; e < nu * nx + nu; e += 32) {
      if (e < nu * nx) Ks[(size_t)(t - 1) * nu * nx + e] = s_K[e];
      else ks[(size_t)(t - 1) * nu + (e - nu * nx)] = s_k[e - nu * nx];
    }
    if (lane < nV) do_V(lane, abV, Vnn, vnn);
    for (int e = lane + 32; e < nV; e += 32) do_V(e, e < nx * nx ? s_tab_x[e] : 0, Vnn, vnn);
    __syncwarp();
    double *tp = Vn; Vn = Vnn; Vnn = tp;
    tp = vn; vn = vnn; vnn = tp;
  }
  s_lin = lin_acc; s_quad = quad_acc;
  if (lane == 0) { *lin_out = s_lin; *quad_out = s_quad; }
}

template <bool RES>
__global__ void __launch_bounds__(NT) ilqr_kernel(const IlqrParams P) {
  extern __shared__ __align__(16) double sm[];
  const long long t_entry = clock64();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const IlqrNet &net = P.net;
  const int nx = P.nx, nu = P.nu, n = nx + nu, H = P.H, LS = P.ls_max_iter, mw = net.max_width;
  // line-search groups: G warps per alpha (2 when they fit), group j = warps [j*G, (j+1)*G)
  const int G = (2 * LS <= NWARPS) ? 2 : 1;
  const int gthr = 32 * G;
  const int grp = warp / G, gl = tid - grp * gthr;
  // ---- shared-memory carve (doubles, then the index tables)
  double *s_Ct = sm;                 // n*n   dt*blkdiag(Q+Q^T, R+R^T)           ilqr.py:170-171
  double *s_Fs = s_Ct + n * n;       // nx*nx F+F^T                             cost.py:208-211
  double *s_V0 = s_Fs + nx * nx;     // nx*nx value Hessian (ping)
  double *s_V1 = s_V0 + nx * nx;     // nx*nx (pong)
  double *s_v0 = s_V1 + nx * nx;     // nx
  double *s_v1 = s_v0 + nx;          // nx
  double *s_T = s_v1 + nx;           // nx*n  Vn @ J
  double *s_Qt = s_T + nx * n;       // n*n
  double *s_qt = s_Qt + n * n;       // n
  double *s_K = s_qt + n;            // nu*nx
  double *s_k = s_K + nu * nx;       // nu
  double *s_LU = s_k + nu;           // nu*nu  LU factors of Quu
  double *s_red = s_LU + nu * nu;    // NWARPS
  double *s_obj = s_red + NWARPS;    // LS + 4
  double *s_cst = sm + (((size_t)(s_obj + LS + 4 - sm) + 1) & ~(size_t)1);   // Q | R | F | goal | goalF | umin | umax | alphas (16-byte aligned from here on)
  double *s_var = s_cst + ((cst_doubles(nx, nu, LS) + 1) & ~(size_t)1);
  const TrajLayout tl(H, nx, nu, LS);
  const size_t jpw = jac_per_warp(mw, n);
  const size_t net_d = ((size_t)net.total + 1) & ~(size_t)1;
  // resident: network | trajectory block, then (both modes) the phase scratch: the line search's activations and
  // K-quarter partials, which the Jacobian refresh's per-warp panels overlay when resident (the phases never overlap);
  // all derived from the shared array
  const double *nb = RES ? s_var : P.net_blob;
  double *traj = RES ? s_var + net_d : P.traj;
  double *s_h = s_var + (RES ? net_d + tl.total : 0);
  const size_t lss = P.ls_mma ? (size_t)P.ls_doubles : ls_scratch(mw, LS);
  const size_t phase_d = RES ? (lss > JAC_WARPS * jpw ? lss : JAC_WARPS * jpw) : lss;
  double *jac_wk = (RES ? s_h : P.jac_work) + (size_t)(warp < JAC_WARPS ? warp : 0) * jpw;
  int *s_tab_n = reinterpret_cast<int *>(s_h + phase_d);                          // e -> (e / n) << 16 | e % n
  int *s_tab_x = s_tab_n + n * n;                                                                    // e -> (e / nx) << 16 | e % nx
  __shared__ int s_flag[4];          // [0]=line search failed, [1]=used idx, [2]=refresh jac
  __shared__ LsArgs s_ls;
  __shared__ int s_piv[MAX_NU];
  __shared__ double s_lin, s_quad;

  double *states = traj + tl.states, *ctrls = traj + tl.ctrls, *Ks = traj + tl.Ks, *ks = traj + tl.ks,
         *Jacs = traj + tl.Jacs, *ls_states = traj + tl.ls_states, *ls_ctrls = traj + tl.ls_ctrls,
         *step_cost = traj + tl.step_cost;
  const double *c_Q = s_cst, *c_R = c_Q + nx * nx, *c_F = c_R + nu * nu, *c_goal = c_F + nx * nx, *c_goalF = c_goal + nx,
               *c_umin = c_goalF + nx, *c_umax = c_umin + nu, *c_alphas = c_umax + nu;
  const double *xu_mean = nb + net.xu_mean, *xu_std = nb + net.xu_std, *dy_mean = nb + net.dy_mean, *dy_std = nb + net.dy_std;

  if (RES) {
    double *dst = s_var;
    for (int t = tid; t < net.total; t += NT) dst[t] = P.net_blob[t];
  }
  if (RES && P.ls_mma && tid == 0) {                   // arguments of the out-of-line line search (shared addresses)
    LsArgs &A = s_ls;
    A.nx = nx; A.nu = nu; A.H = H; A.S = P.ls_S; A.mwp = P.ls_mwp; A.L = net.n_layers; A.act = net.act; A.bounded = P.bounded;
    A.profile = P.ls_profile;
    const uint32_t a_nb = sh32(nb), a_scr = sh32(s_h);
    for (int l = 0; l < net.n_layers; ++l) {
      A.KS[l] = P.ls_kp[l] >> 2; A.MT[l] = P.ls_mt[l]; A.N[l] = net.dims[l + 1]; A.Kin[l] = net.dims[l]; A.ws[l] = net.wstride[l];
      A.a_W[l] = a_nb + (uint32_t)net.woff[l] * 8u; A.a_B[l] = a_nb + (uint32_t)net.boff[l] * 8u;
      A.a_Wf[l] = a_scr + (uint32_t)(2 * P.ls_mwp * P.ls_S + 8 * 64 + P.ls_wf[l]) * 8u;
    }
    A.a_xu_mean = a_nb + (uint32_t)net.xu_mean * 8u; A.a_xu_std = a_nb + (uint32_t)net.xu_std * 8u;
    A.a_dy_mean = a_nb + (uint32_t)net.dy_mean * 8u; A.a_dy_std = a_nb + (uint32_t)net.dy_std * 8u;
    A.a_scr = a_scr; A.a_states = sh32(states); A.a_ctrls = sh32(ctrls); A.a_Ks = sh32(Ks); A.a_ks = sh32(ks);
    A.a_ls_states = sh32(ls_states); A.a_ls_ctrls = sh32(ls_ctrls); A.a_alphas = sh32(c_alphas); A.a_umin = sh32(c_umin);
    A.a_umax = sh32(c_umax);
    A.x0 = P.x0; A.prof = P.prof;
    A.a_W0s = a_nb + (uint32_t)net.w0s * 8u;
    const uint32_t a_after = a_scr + (uint32_t)(2 * P.ls_mwp * P.ls_S + 8 * 64 + P.ls_wf_total) * 8u;
    A.SP = P.jac_SP; A.ncols = 8 * n;
    A.a_G = a_after;
    A.a_PA = A.a_G + (uint32_t)((net.n_layers - 1) * P.ls_mwp * 8) * 8u;
    A.a_PB = A.a_PA + (uint32_t)(P.ls_mwp * P.jac_SP) * 8u;
    A.a_tab = A.a_PB + (uint32_t)(P.ls_mwp * P.jac_SP) * 8u;
    A.a_Jacs = sh32(Jacs);
    A.wf_persistent = P.jac_mma;
  }
  for (int t = tid; t < (int)cst_doubles(nx, nu, LS); t += NT) s_cst[t] = P.cst[t];
  for (int t = tid; t < n * n; t += NT) s_tab_n[t] = ((t / n) << 16) | (t % n);
  for (int t = tid; t < nx * nx; t += NT) s_tab_x[t] = ((t / nx) << 16) | (t % nx);
  __syncthreads();
  for (int t = tid; t < n * n; t += NT) {
    const int r = t / n, c = t - r * n;
    double v = 0.0;
    if (r < nx && c < nx) v = P.dt * (c_Q[r * nx + c] + c_Q[c * nx + r]);
    else if (r >= nx && c >= nx) v = P.dt * (c_R[(r - nx) * nu + (c - nx)] + c_R[(c - nx) * nu + (r - nx)]);
    s_Ct[t] = v;
  }
  for (int t = tid; t < nx * nx; t += NT) {
    const int r = t / nx, c = t - r * nx;
    s_Fs[t] = c_F[r * nx + c] + c_F[c * nx + r];
  }
  for (int t = tid; t < nx; t += NT) states[t] = P.x0[t];
  for (int t = tid; t < H * nu; t += NT) ctrls[t] = P.uguess ? P.uguess[t] : 0.0;
  for (int t = tid; t < P.max_iter; t += NT) P.alpha_idx[t] = -1;
  if (tid < 8) P.prof[8 + tid] = 0ull;
  __syncthreads();

  // per-step cost table for trajectory (xs (H+1,nx), us (H,nu)): c[i] = dt*(obs+ctrl), c[H] = terminal   (ilqr.py:124-129)
  auto step_cost_of = [&](const double *xs, const double *us, int i) -> double {
    if (i < H) return P.dt * (quad_form(c_Q, xs + (size_t)i * nx, c_goal, nx) + quad_form(c_R, us + (size_t)i * nu, nullptr, nu));
    return quad_form(c_F, xs + (size_t)H * nx, c_goalF, nx);
  };

  if (RES && P.jac_mma) {                                // tensor-core phases: zero their scratch, build the image and the table once
    for (int t = tid; t < (int)phase_d; t += NT) s_h[t] = 0.0;
    __syncthreads();
    ls_build_wf(&s_ls, NT);
    for (int t = tid; t < 8 * n; t += NT) asm volatile("st.shared.u32 [%0], %1;" ::"r"(s_ls.a_tab + (uint32_t)t * 4u), "r"(t / n) : "memory");
    __syncthreads();
  }
  // ---- initial rollout (ilqr.py:141-147) by line-search group 0; Jacobians are evaluated in one batch afterwards
  if (grp == 0) {
    double *h0 = s_h, *h1 = s_h + mw;
    for (int i = 0; i < H; ++i) {
      const double *xi = states + (size_t)i * nx;
      for (int j = gl; j < n; j += gthr) {
        const double v = j < nx ? xi[j] : ctrls[(size_t)i * nu + (j - nx)];
        h0[j] = (v - xu_mean[j]) / xu_std[j];
      }
      group_sync(1, gthr);
      const double *out = group_forward(net, nb, h0, h1, gl, gthr, 1);
      for (int j = gl; j < nx; j += gthr) states[(size_t)(i + 1) * nx + j] = xi[j] + (out[j] * dy_std[j] + dy_mean[j]);
      group_sync(1, gthr);
    }
  }
  __syncthreads();
  if (RES && P.jac_mma) jac_batch_mma(&s_ls, sh32(states), sh32(ctrls));
  else jac_batch(P, nb, states, ctrls, Jacs, jac_wk, warp, lane);
  for (int i = tid; i <= H; i += NT) step_cost[i] = step_cost_of(states, ctrls, i);
  __syncthreads();
  double obj = 0.0;       // every thread tracks the same scalars (uniform control flow)
  for (int i = 0; i <= H; ++i) obj += step_cost[i];     // sequential like eval_obj, ilqr.py:124-129
  __syncthreads();

  long long t_mark = clock64();
  unsigned long long cyc[6] = {0, 0, 0, 0, 0, 0};
  auto lap = [&](int slot) { const long long now = clock64(); cyc[slot] += (unsigned long long)(now - t_mark); t_mark = now; };
  cyc[0] = (unsigned long long)(t_mark - t_entry);
  int converged = 0, n_iter = 0, ls_fail = 0;
  for (int itr = 0; itr < P.max_iter; ++itr) {
    n_iter = itr + 1;
    // ---- backward pass (ilqr.py:159-187): one warp (backward_pass)
    if (warp == 0) {
      BackwardSmem bs{s_Ct, s_Fs, s_V0, s_V1, s_v0, s_v1, s_T, s_Qt, s_qt, s_K, s_k, s_LU, s_tab_n, s_tab_x, s_piv};
      if (nx == 4 && nu == 1) backward_pass<4, 1>(P, bs, states, ctrls, Jacs, Ks, ks, c_goal, &s_lin, &s_quad, lane);
      else backward_pass<0, 0>(P, bs, states, ctrls, Jacs, Ks, ks, c_goal, &s_lin, &s_quad, lane);
    }
    __syncthreads();
    lap(1);
    const double lin_cost_reduce = s_lin, quad_cost_reduce = s_quad;
    double ksq = 0.0;
    for (int t = tid; t < H * nu; t += NT) ksq += ks[t] * ks[t];
    const double ks_norm = sqrt(block_sum(ksq, s_red, tid));

    // ---- line-search rollouts (ilqr.py:190-205) and the objective of every step size (per-step costs in parallel, then
    //      a sequential sum per step size)
    auto objectives = [&](int a0, int na) {
      for (int t = tid; t < na * (H + 1); t += NT) {
        const int j = a0 + t / (H + 1), i = t % (H + 1);
        step_cost[(H + 1) + j * (H + 1) + i] = step_cost_of(ls_states + (size_t)j * (H + 1) * nx, ls_ctrls + (size_t)j * H * nu, i);
      }
      __syncthreads();
      if (tid < na) {
        double o = 0.0;
        for (int i = 0; i <= H; ++i) o += step_cost[(H + 1) + (a0 + tid) * (H + 1) + i];
        s_obj[a0 + tid] = o;
      }
      __syncthreads();
    };
    if (P.ls_mma) {
      // eight step sizes at a time; the acceptance rule below walks them in order and stops at the first one it takes
      // (or at once when the step is tiny), so a later tile is rolled out only if the rule would get that far
      for (int a0 = 0; a0 < LS; a0 += 8) {
        const int na = LS - a0 < 8 ? LS - a0 : 8;
        if (warp < LS_WARPS)
        {
          if (nx == 4) ls_rollouts_mma<4>(&s_ls, a0, na);
          else ls_rollouts_mma<0>(&s_ls, a0, na);
        }
        __syncthreads();
        lap(2);
        objectives(a0, na);
        if (tid == 0) {
          int stop = 0;
          for (int l = a0; l < a0 + na && !stop; ++l) {
            const double alpha = c_alphas[l];
            const double expect = alpha * lin_cost_reduce + alpha * alpha * quad_cost_reduce / 2;
            if ((obj - s_obj[l]) / (-expect) > P.ls_cost_threshold || ks_norm < P.u_threshold) stop = 1;
          }
          s_flag[3] = stop;
        }
        __syncthreads();
        lap(3);
        if (s_flag[3]) break;
      }
    } else {
      if (warp < LS_WARPS) {
        double *part = s_h + (size_t)2 * mw * ls_cols(LS);
        if (LS <= 10) ls_rollouts<10, 1>(P, nb, s_h, part, states, ctrls, Ks, ks, ls_states, ls_ctrls, c_alphas, c_umin, c_umax, tid);
        else ls_rollouts<10, 2>(P, nb, s_h, part, states, ctrls, Ks, ks, ls_states, ls_ctrls, c_alphas, c_umin, c_umax, tid);
      }
      __syncthreads();
      lap(2);
      objectives(0, LS);
    }
    // ---- backtracking acceptance (ilqr.py:208-238), thread 0 decides
    if (tid == 0) {
      int best_idx = -1, used = -1, have_best = 0;
      double best_obj = INFINITY, new_obj = 0.0;
      for (int l = 0; l < LS; ++l) {
        used = l;
        new_obj = s_obj[l];
        const double alpha = c_alphas[l];
        const double expect = alpha * lin_cost_reduce + alpha * alpha * quad_cost_reduce / 2;
        if ((obj - new_obj) / (-expect) > P.ls_cost_threshold) { best_obj = new_obj; best_idx = l; have_best = 1; break; }
        if (new_obj < best_obj) { best_obj = new_obj; best_idx = l; have_best = 1; }
        if (ks_norm < P.u_threshold) break;
      }
      int ls_success = 0;
      if (best_obj < obj || ks_norm < P.u_threshold) {
        // NB: the reference indexes ls_*[best_alpha_idx] here; with no best it would raise -- treat as failure
        if (have_best) { ls_success = 1; used = best_idx; new_obj = s_obj[best_idx]; }
      }
      int fail = ((!ls_success && new_obj > obj + 1e-3) || !have_best) ? 1 : 0;
      s_flag[0] = fail;
      s_flag[1] = used;
      s_flag[2] = ls_success;
      s_obj[LS] = new_obj;
    }
    __syncthreads();
    lap(3);
    if (s_flag[0]) { ls_fail = 1; break; }
    const int used = s_flag[1];
    const double *nxs = ls_states + (size_t)used * (H + 1) * nx, *nus = ls_ctrls + (size_t)used * H * nu;
    if (s_flag[2]) {                                       // ilqr.py:232 (stale Jacobians otherwise, as in the reference)
      if (RES && P.jac_mma) jac_batch_mma(&s_ls, sh32(nxs), sh32(nus));
      else jac_batch(P, nb, nxs, nus, Jacs, jac_wk, warp, lane);
    }
    lap(4);
    if (tid == 0) P.alpha_idx[itr] = used;
    double dsq = 0.0;
    for (int t = tid; t < H * nu; t += NT) { const double d = nus[t] - ctrls[t]; dsq += d * d; }
    const double du_norm = sqrt(block_sum(dsq, s_red, tid));   // ilqr.py:246
    if (du_norm < P.u_threshold) converged = 1;
    for (int t = tid; t < (H + 1) * nx; t += NT) states[t] = nxs[t];
    for (int t = tid; t < H * nu; t += NT) ctrls[t] = nus[t];
    obj = s_obj[LS];
    __syncthreads();
    lap(3);
    if (converged) break;
  }
  __syncthreads();
  for (int t = tid; t < (H + 1) * nx; t += NT) P.states[t] = states[t];
  for (int t = tid; t < H * nu; t += NT) { P.ctrls[t] = ctrls[t]; P.ks[t] = ks[t]; }
  for (int t = tid; t < H * nu * nx; t += NT) P.Ks[t] = Ks[t];
  if (tid == 0) {
    P.info[0] = converged; P.info[1] = n_iter; P.info[2] = ls_fail;
    lap(5);
    for (int q = 0; q < 6; ++q) P.prof[q] = cyc[q];
    P.prof[6] = (unsigned long long)(clock64() - t_entry);
    P.prof[7] = (unsigned long long)n_iter;
  }
}

}  // namespace

struct ampc_ilqr {
  ampc_ilqr_cfg cfg;
  IlqrParams P;
  int device = 0;
  bool resident = false;
  double *d_net = nullptr;    // network blob (row-major padded weights, biases, normalisers)
  double *d_work = nullptr;   // everything else
  int *d_int = nullptr;
  unsigned long long *d_prof = nullptr;
  size_t smem = 0;
  size_t o_x0 = 0, o_ug = 0;
};

static void launch_ilqr(const ampc_ilqr *h, const IlqrParams &P, cudaStream_t s) {
  if (h->resident) ilqr_kernel<true><<<1, NT, h->smem, s>>>(P);
  else ilqr_kernel<false><<<1, NT, h->smem, s>>>(P);
  ampc_count_launch();
}

extern "C" int ampc_ilqr_create(ampc_ilqr **out, const ampc_ilqr_cfg *cfg, const ampc_mlp_desc *mlp,
                                const ampc_quad_cost *cost) {
  AMPC_REQUIRE(out && cfg && mlp && cost, AMPC_ERR_INVALID, "null argument");
  *out = nullptr;
  AMPC_REQUIRE(cfg->H >= 1 && cfg->nx >= 1 && cfg->nu >= 1 && cfg->nu <= MAX_NU && cfg->nx + cfg->nu <= AMPC_MAX_WIDTH,
               AMPC_ERR_INVALID, "bad iLQR dims H=%d nx=%d nu=%d (nu <= %d)", cfg->H, cfg->nx, cfg->nu, MAX_NU);
  AMPC_REQUIRE(cfg->max_iter >= 1 && cfg->ls_max_iter >= 1 && cfg->ls_max_iter <= MAX_LS, AMPC_ERR_INVALID,
               "bad iteration limits (1 <= ls_max_iter <= %d)", MAX_LS);
  AMPC_REQUIRE(mlp->n_layers >= 2 && mlp->n_layers <= AMPC_MAX_LAYERS, AMPC_ERR_UNSUPPORTED,
               "MLP must have 1..%d hidden layers", AMPC_MAX_LAYERS - 1);
  AMPC_REQUIRE(mlp->dims[0] == cfg->nx + cfg->nu && mlp->dims[mlp->n_layers] == cfg->nx, AMPC_ERR_INVALID,
               "MLP dims do not match nx+nu -> nx");
  AMPC_REQUIRE(mlp->act >= 0 && mlp->act <= 3, AMPC_ERR_UNSUPPORTED, "unknown activation %d", mlp->act);
  for (int l = 0; l <= mlp->n_layers; ++l)
    AMPC_REQUIRE(mlp->dims[l] >= 1 && mlp->dims[l] <= AMPC_MAX_WIDTH, AMPC_ERR_UNSUPPORTED, "layer width %d", mlp->dims[l]);
  for (int j = 0; j < cfg->nx + cfg->nu; ++j) AMPC_REQUIRE(mlp->xu_std[j] != 0.0, AMPC_ERR_INVALID, "xu_std[%d] == 0", j);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  AMPC_REQUIRE(ce == cudaSuccess && ndev > 0, AMPC_ERR_CUDA, "no CUDA device: libampc_b200 has no CPU fallback (%s)",
               cudaGetErrorString(ce));
  AMPC_REQUIRE(cfg->device >= 0 && cfg->device < ndev, AMPC_ERR_INVALID, "device %d of %d", cfg->device, ndev);
  AMPC_CUDA_CHECK(cudaSetDevice(cfg->device));
  const int H = cfg->H, nx = cfg->nx, nu = cfg->nu, n = nx + nu, LS = cfg->ls_max_iter;
  ampc_ilqr *h = new ampc_ilqr();
  h->cfg = *cfg;
  h->device = cfg->device;
  IlqrParams &P = h->P;
  memset(&P, 0, sizeof(P));
  // ---- network blob: W_l row-major (out, in) with an odd row stride, then b_l, then the normalisers
  IlqrNet &net = P.net;
  net.n_layers = mlp->n_layers;
  net.act = mlp->act;
  int off = 0;
  for (int l = 0; l <= mlp->n_layers; ++l) {
    net.dims[l] = mlp->dims[l];
    if (mlp->dims[l] > net.max_width) net.max_width = mlp->dims[l];
  }
  for (int l = 0; l < mlp->n_layers; ++l) {
    net.wstride[l] = mlp->dims[l] | 1;
    net.woff[l] = off; off += net.wstride[l] * mlp->dims[l + 1]; off = (off + 1) & ~1;
    net.boff[l] = off; off += mlp->dims[l + 1]; off = (off + 1) & ~1;
  }
  net.w0s = off; off += (mlp->dims[0] * mlp->dims[1] + 1) & ~1;
  net.xu_mean = off; off += (n + 1) & ~1;
  net.xu_std = off; off += (n + 1) & ~1;
  net.dy_mean = off; off += (nx + 1) & ~1;
  net.dy_std = off; off += (nx + 1) & ~1;
  net.total = off;
  std::vector<double> hb(off, 0.0);
  for (int l = 0; l < mlp->n_layers; ++l) {
    const int Kin = mlp->dims[l], N = mlp->dims[l + 1];
    for (int j = 0; j < N; ++j) {
      for (int k = 0; k < Kin; ++k) hb[net.woff[l] + (size_t)j * net.wstride[l] + k] = mlp->W[l][(size_t)j * Kin + k];
      hb[net.boff[l] + j] = mlp->b[l][j];
    }
  }
  for (int j = 0; j < mlp->dims[1]; ++j)
    for (int c = 0; c < n; ++c) hb[net.w0s + (size_t)j * n + c] = mlp->W[0][(size_t)j * n + c] / mlp->xu_std[c];
  for (int j = 0; j < n; ++j) { hb[net.xu_mean + j] = mlp->xu_mean[j]; hb[net.xu_std + j] = mlp->xu_std[j]; }
  for (int j = 0; j < nx; ++j) { hb[net.dy_mean + j] = mlp->dy_mean[j]; hb[net.dy_std + j] = mlp->dy_std[j]; }
  const int mw = net.max_width;
  P.H = H; P.nx = nx; P.nu = nu; P.bounded = cfg->bounded; P.max_iter = cfg->max_iter; P.ls_max_iter = LS;
  P.dt = cfg->dt; P.ls_discount = cfg->ls_discount; P.ls_cost_threshold = cfg->ls_cost_threshold;
  P.u_threshold = cfg->u_threshold;
  // ---- global work area: constants | x0 | uguess | outputs | trajectory block | Jacobian scratch
  const TrajLayout tl(H, nx, nu, LS);
  const size_t jpw = jac_per_warp(mw, n);
  size_t woff = 0;
  auto take = [&](size_t cnt) { size_t o = woff; woff += (cnt + 1) & ~(size_t)1; return o; };
  const size_t o_cst = take(cst_doubles(nx, nu, LS));
  h->o_x0 = take(nx); h->o_ug = take((size_t)H * nu);
  const size_t o_st = take((size_t)(H + 1) * nx), o_ct = take((size_t)H * nu), o_Ks = take((size_t)H * nu * nx), o_ks = take((size_t)H * nu);
  const size_t o_traj = take(tl.total), o_jac = take((size_t)JAC_WARPS * jpw + 8);   // + slack: the last column strip may read past a panel
  std::vector<double> hc(cst_doubles(nx, nu, LS), 0.0);
  {
    double *q = hc.data();
    memcpy(q, cost->Q, sizeof(double) * nx * nx); q += nx * nx;
    memcpy(q, cost->R, sizeof(double) * nu * nu); q += nu * nu;
    memcpy(q, cost->F, sizeof(double) * nx * nx); q += nx * nx;
    memcpy(q, cost->goal, sizeof(double) * nx); q += nx;
    memcpy(q, cost->goal_term ? cost->goal_term : cost->goal, sizeof(double) * nx); q += nx;
    memcpy(q, cost->umin, sizeof(double) * nu); q += nu;
    memcpy(q, cost->umax, sizeof(double) * nu); q += nu;
    for (int i = 0; i < LS; ++i) q[i] = pow(cfg->ls_discount, (double)i);   // ls_discount**i, ilqr.py:196
  }
  cudaError_t e = cudaMalloc(&h->d_net, hb.size() * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_net, hb.data(), hb.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_work, woff * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(h->d_work, 0, woff * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_work + o_cst, hc.data(), hc.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_int, (3 + cfg->max_iter) * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_prof, 16 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(h->d_prof, 0, 16 * sizeof(unsigned long long));
  double *w = h->d_work;
  P.net_blob = h->d_net;
  P.cst = w + o_cst;
  P.x0 = w + h->o_x0; P.uguess = nullptr;
  P.states = w + o_st; P.ctrls = w + o_ct; P.Ks = w + o_Ks; P.ks = w + o_ks;
  P.traj = w + o_traj; P.jac_work = w + o_jac;
  P.info = h->d_int; P.alpha_idx = h->d_int + 3; P.prof = h->d_prof;
  {
    // shared memory: the small matrices, the cost constants, the per-group activations and the index tables always;
    // network + trajectory block + Jacobian scratch when they all fit next to them
    size_t fixed = (size_t)n * n * 2 + (size_t)nx * nx * 3 + 2 * nx + (size_t)nx * n + n + (size_t)nu * nx + nu +
                   (size_t)nu * nu + NWARPS + LS + 4 + 1 + ((cst_doubles(nx, nu, LS) + 1) & ~(size_t)1);
    const size_t tabs = ((size_t)n * n + (size_t)nx * nx + 1) / 2 + 1;       // ints, in doubles
    // line search on mma.m8n8k4.f64 fragments: activations [padded rows][S] x 2 + the fragment-ordered weight image
    size_t lss_mma = 0, jac_mma_doubles = 0;
    {
      int mwp = 0, off = 0;
      P.ls_S = 20;                                       // = 4 (mod 16) doubles: conflict-free operand loads
      for (int l = 0; l < net.n_layers; ++l) {
        P.ls_kp[l] = (net.dims[l] + 15) & ~15;           // k-steps come in groups of four (one per partial accumulator)
        P.ls_mt[l] = (net.dims[l + 1] + 7) / 8;
        P.ls_wf[l] = off;
        off += P.ls_mt[l] * 8 * P.ls_kp[l];
        mwp = std::max(mwp, std::max(P.ls_kp[l], P.ls_mt[l] * 8));
      }
      P.ls_mwp = mwp;
      P.ls_wf_total = off;
      lss_mma = (size_t)2 * mwp * P.ls_S + 8 * 64 + off;  // + the K-split partials of the output layer
      // Jacobian refresh on the same fragments: + act' per hidden layer, two panels of 8 (nx + nu) columns, column table
      const int ncols = 8 * n, ncp = (ncols + 7) & ~7;
      P.jac_SP = ncp + ((ncp & 7) == 4 ? 0 : 4);         // = 4 (mod 8): conflict-free operand loads
      jac_mma_doubles = lss_mma + (size_t)(net.n_layers - 1) * mwp * 8 + (size_t)2 * mwp * P.jac_SP + (ncols + 1) / 2 + 2;
      P.ls_doubles = (int)lss_mma;
    }
    const size_t jacs = (size_t)JAC_WARPS * jpw, lss_old = ls_scratch(mw, LS);
    size_t fixed_probe = (size_t)n * n * 2 + (size_t)nx * nx * 3 + 2 * nx + (size_t)nx * n + n + (size_t)nu * nx + nu +
                         (size_t)nu * nu + NWARPS + LS + 4 + 1 + ((cst_doubles(nx, nu, LS) + 1) & ~(size_t)1) +
                         ((size_t)n * n + (size_t)nx * nx + 1) / 2 + 1;
    int optin_probe = 0;
    cudaDeviceGetAttribute(&optin_probe, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
    const size_t cap_probe = (size_t)optin_probe > 2048 ? ((size_t)optin_probe - 2048) / sizeof(double) : 0;
    {
      // taken in resident mode only (it addresses every array as shared memory), when everything still fits with its
      // phase scratch
      const size_t base = (((size_t)net.total + 1) & ~(size_t)1) + tl.total;
      const bool res_old = fixed_probe + base + std::max(jacs, lss_old) <= cap_probe;
      const bool res_mma = fixed_probe + base + std::max(jacs, lss_mma) <= cap_probe;
      (void)res_old;
      P.ls_profile = getenv("AMPC_ILQR_LS_PROFILE") ? 1 : 0;
      P.ls_mma = (!getenv("AMPC_ILQR_NO_MMA") && !getenv("AMPC_ILQR_NO_SMEM") && res_mma && 8 * n <= NT / 2) ? 1 : 0;
    }
    // the tensor-core Jacobian refresh lives in the phase scratch the per-warp routine would use
    P.jac_mma = (P.ls_mma && !getenv("AMPC_ILQR_NO_JAC_MMA") && jac_mma_doubles <= std::max(jacs, lss_mma)) ? 1 : 0;
    const size_t lss = P.ls_mma ? lss_mma : lss_old;
    const size_t var = (((size_t)net.total + 1) & ~(size_t)1) + tl.total + (jacs > lss ? jacs : lss);
    fixed += tabs;
    int max_optin = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
    const size_t cap = (size_t)max_optin > 2048 ? ((size_t)max_optin - 2048) / sizeof(double) : 0;
    h->resident = (fixed + var <= cap) && !getenv("AMPC_ILQR_NO_SMEM");
    const size_t total = fixed + (h->resident ? var : lss);
    h->smem = total * sizeof(double);
    if (e == cudaSuccess && total > cap) {
      ampc_set_error("iLQR: %zu B of shared memory needed", h->smem);
      cudaFree(h->d_net); cudaFree(h->d_work); cudaFree(h->d_int); cudaFree(h->d_prof);
      delete h;
      return AMPC_ERR_UNSUPPORTED;
    }
  }
  if (e == cudaSuccess)
    e = h->resident ? ampc_raise_smem_limit((const void *)ilqr_kernel<true>, h->smem)
                    : ampc_raise_smem_limit((const void *)ilqr_kernel<false>, h->smem);
  if (e != cudaSuccess) {
    ampc_set_error("iLQR create: %s", cudaGetErrorString(e));
    cudaFree(h->d_net); cudaFree(h->d_work); cudaFree(h->d_int); cudaFree(h->d_prof);
    delete h;
    return AMPC_ERR_CUDA;
  }
  *out = h;
  return AMPC_OK;
}

extern "C" int ampc_ilqr_destroy(ampc_ilqr *h) {
  if (!h) return AMPC_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_net); cudaFree(h->d_work); cudaFree(h->d_int); cudaFree(h->d_prof);
  delete h;
  return AMPC_OK;
}

// Debug tap, no reference counterpart: SM cycles thread 0 of the last solve spent per phase (see IlqrParams::prof).
extern "C" int ampc_ilqr_debug_profile(ampc_ilqr *h, unsigned long long *out8) {
  AMPC_REQUIRE(h && out8, AMPC_ERR_INVALID, "null argument");
  AMPC_CUDA_CHECK(cudaSetDevice(h->device));
  AMPC_CUDA_CHECK(cudaDeviceSynchronize());
  if (getenv("AMPC_ILQR_LS_PROFILE")) {   // the line-search phase split (thread 0): controls, main loops, barrier, combine, barrier, update
    unsigned long long v[16];
    AMPC_CUDA_CHECK(cudaMemcpy(v, h->d_prof, sizeof(v), cudaMemcpyDeviceToHost));
    fprintf(stderr, "ilqr line search (cycles of thread 0; ls_rollouts: controls, main loops, barrier, combine, barrier, update; ls_rollouts_mma: controls, layer 0, output layer, other layers, barrier, update): %llu %llu %llu %llu %llu %llu\n",
            v[8], v[9], v[10], v[11], v[12], v[13]);
  }
  AMPC_CUDA_CHECK(cudaMemcpy(out8, h->d_prof, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return AMPC_OK;
}

// Re-runs the solve on whatever x0 / uguess the last solve_host left on the device, asynchronously on `stream`: no
// host copies.  Used to time the kernel alone (bench.py --workload c4); outputs stay on the device.
extern "C" int ampc_ilqr_launch(ampc_ilqr *h, void *stream) {
  AMPC_REQUIRE(h, AMPC_ERR_INVALID, "null handle");
  AMPC_CUDA_CHECK(cudaSetDevice(h->device));
  launch_ilqr(h, h->P, (cudaStream_t)stream);
  AMPC_CUDA_CHECK(cudaGetLastError());
  return AMPC_OK;
}

extern "C" int ampc_ilqr_solve_host(ampc_ilqr *h, const double *x0, const double *uguess, double *states,
                                    double *ctrls, double *Ks, double *ks, int32_t *info, int32_t *alpha_idx) {
  AMPC_REQUIRE(h && x0 && states && ctrls && Ks && ks && info, AMPC_ERR_INVALID, "null argument");
  AMPC_CUDA_CHECK(cudaSetDevice(h->device));
  const int H = h->cfg.H, nx = h->cfg.nx, nu = h->cfg.nu;
  IlqrParams P = h->P;
  AMPC_CUDA_CHECK(cudaMemcpy(h->d_work + h->o_x0, x0, nx * sizeof(double), cudaMemcpyHostToDevice));
  if (uguess) {
    AMPC_CUDA_CHECK(cudaMemcpy(h->d_work + h->o_ug, uguess, (size_t)H * nu * sizeof(double), cudaMemcpyHostToDevice));
    P.uguess = h->d_work + h->o_ug;
  }
  launch_ilqr(h, P, nullptr);
  AMPC_CUDA_CHECK(cudaGetLastError());
  AMPC_CUDA_CHECK(cudaMemcpy(states, P.states, (size_t)(H + 1) * nx * sizeof(double), cudaMemcpyDeviceToHost));
  AMPC_CUDA_CHECK(cudaMemcpy(ctrls, P.ctrls, (size_t)H * nu * sizeof(double), cudaMemcpyDeviceToHost));
  AMPC_CUDA_CHECK(cudaMemcpy(Ks, P.Ks, (size_t)H * nu * nx * sizeof(double), cudaMemcpyDeviceToHost));
  AMPC_CUDA_CHECK(cudaMemcpy(ks, P.ks, (size_t)H * nu * sizeof(double), cudaMemcpyDeviceToHost));
  AMPC_CUDA_CHECK(cudaMemcpy(info, P.info, 3 * sizeof(int), cudaMemcpyDeviceToHost));
  if (alpha_idx)
    AMPC_CUDA_CHECK(cudaMemcpy(alpha_idx, P.alpha_idx, h->cfg.max_iter * sizeof(int), cudaMemcpyDeviceToHost));
  return AMPC_OK;
}
