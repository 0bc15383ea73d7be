// Iterative LQR solve on the device, float64, one persistent CTA per solve.
//
// Replaces autompc.control.ilqr.IterativeLQR.compute_ilqr_default
// (autompc/control/ilqr.py:100-265) for MLP dynamics + QuadCost:
//   init rollout (:141-149)  ->  up to max_iter x [ backward Riccati (:159-187),
//   batched ls_max_iter-alpha line search (:197-225), Jacobian refresh (:226-234),
//   stopping rule (:235-261) ].
// The problem is tiny and strictly sequential in H (state recursion forward,
// value recursion backward), so the whole solve is ONE launch: no host round
// trips between the ~50 x (H Riccati steps + H line-search steps) stages the
// reference walks in Python.  float64 keeps the integer outputs (adopted
// line-search index per iteration, iteration count, converged) equal to the
// reference's.  The small (nx+nu)^2 value / gain matrices live in shared memory;
// trajectories, gains and the per-step Jacobian panels live in a global scratch
// that stays in L1/L2.
#include <vector>

#include "ampc_common.cuh"
#include "mlp_f64.cuh"

namespace {

constexpr int NT = 640;            // 20 warps: warp 0 runs the Riccati recursion, line-search rollouts use G warps each
constexpr int NWARPS = NT / 32;
constexpr int MAX_NU = 16;
constexpr int MAX_LS = 20;

struct IlqrParams {
  AmpcMlpF64 net;
  int H, nx, nu, bounded, max_iter, ls_max_iter;
  double dt, ls_discount, ls_cost_threshold, u_threshold;
  const double *Q, *R, *F, *goal, *goalF, *umin, *umax, *alphas;   // device (goalF: goal of the terminal term)
  const double *x0, *uguess;                       // device (uguess may be null)
  // scratch / outputs (device, global).  With traj_smem != 0 the kernel keeps the trajectories, gains, Jacobians and
  // line-search rollouts in shared memory instead and writes the outputs back once at the end.
  double *states, *ctrls, *Ks, *ks, *Jacs, *ls_states, *ls_ctrls, *step_cost;
  double *hA;                      // global per-warp scratch of the Jacobian refresh (when it does not fit shared memory)
  int *info, *alpha_idx;
  unsigned long long *prof;        // [0..5] cycles of thread 0 in: setup+init rollout, backward passes, line-search rollouts,
                                   // objective + acceptance, Jacobian refreshes, copy-out; [6] = total, [7] = iterations
  int w_smem;                      // != 0: the network's weights and biases are staged into shared memory
  int traj_smem;
  int jac_smem;                    // != 0: the per-warp scratch of the Jacobian refresh is shared memory (else hA)
  size_t net_doubles;              // weights + biases + normalisers as laid out by ampc_mlp_f64_upload
  const double *net_blob;
};

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

// per-step cost table for trajectory (xs (H+1,nx), us (H,nu)): c[i] = dt*(obs+ctrl), c[H] = terminal
__device__ __forceinline__ void step_costs(const IlqrParams &P, const double *xs, const double *us, double *c,
                                           int tid, int stride_thr) {
  for (int i = tid; i <= P.H; i += stride_thr) {
    if (i < P.H)
      c[i] = P.dt * (quad_form(P.Q, xs + (size_t)i * P.nx, P.goal, P.nx) + quad_form(P.R, us + (size_t)i * P.nu, nullptr, P.nu));
    else
      c[i] = quad_form(P.F, xs + (size_t)P.H * P.nx, P.goalF, P.nx);
  }
}

// dot(Wt[:, j], h), four partial sums like ampc_dot_col but with plain loads: Wt may be shared OR global memory
__device__ __forceinline__ double dot_col_any(const double *Wt, int N, int j, const double *h, int Kin) {
  double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
  int k = 0;
#pragma unroll 4
  for (; k + 4 <= Kin; k += 4) {
    p0 = fma(Wt[(size_t)(k + 0) * N + j], h[k + 0], p0);
    p1 = fma(Wt[(size_t)(k + 1) * N + j], h[k + 1], p1);
    p2 = fma(Wt[(size_t)(k + 2) * N + j], h[k + 2], p2);
    p3 = fma(Wt[(size_t)(k + 3) * N + j], h[k + 3], p3);
  }
  for (; k < Kin; ++k) p0 = fma(Wt[(size_t)k * N + j], h[k], p0);
  return (p0 + p1) + (p2 + p3);
}

// barrier over the `nthr` threads of one line-search group (named barrier `id` >= 1), or a warp barrier
__device__ __forceinline__ void group_sync(int id, int nthr) {
  if (nthr == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthr) : "memory");
}

// One MLP forward for ONE sample by a group of `gthr` threads (`gl` = rank inside the group): h0 holds the z-scored
// input, h0/h1 ping-pong (shared memory), returns the buffer with the raw outputs.  Same arithmetic per output as
// ampc_mlp_f64_forward_batch (bit-identical results).
__device__ __forceinline__ const double *group_forward(const AmpcMlpF64 &net, double *h0, double *h1, int gl, int gthr,
                                                       int bar) {
  double *hin = h0, *hout = h1;
  for (int l = 0; l < net.n_layers; ++l) {
    const int Kin = net.dims[l], N = net.dims[l + 1];
    const bool last = (l == net.n_layers - 1);
    for (int j = gl; j < N; j += gthr) {
      const double y = net.b[l][j] + dot_col_any(net.Wt[l], N, j, hin, Kin);
      hout[j] = last ? y : ampc_act<double>(net.act, y);
    }
    group_sync(bar, gthr);
    double *t2 = hin; hin = hout; hout = t2;
  }
  return hin;
}

// Jacobians of x' = x + dy(x,u) at (xs[i], us[i]) for i < H  ->  Jacs (H, nx, nx+nu)   (mlp.py:281-305)
// One WARP per sample (samples warp, warp + NWARPS, ...), warp barriers only.  Forward-mode propagation of the
// [width x nin] panel through the layer stack; per layer a lane owns output rows j = lane, lane + 32, ... and keeps a
// CB-column strip of its row in registers while it walks k, so a weight is loaded once per CB multiply-adds and the
// previous panel's row k is a broadcast load.  `wk` = this warp's scratch: h0, h1, g (mw each) and two panels
// (mw * nin each) -- shared memory when they fit, else a per-warp region of the global scratch.
constexpr int JAC_CB = 8;
__device__ void jac_batch(const IlqrParams &P, const AmpcMlpF64 &net, const double *xs, const double *us, double *Jacs,
                          double *wk, int warp, int lane) {
  const int nx = P.nx, nu = P.nu, nin = nx + nu, H = P.H, mw = net.max_width;
  double *h0 = wk, *h1 = h0 + mw, *g = h1 + mw, *J0 = g + mw, *J1 = J0 + (size_t)mw * nin;
  for (int s = warp; s < H; s += NWARPS) {
    for (int j = lane; j < nin; j += 32) {
      const double v = j < nx ? xs[(size_t)s * nx + j] : us[(size_t)s * nu + (j - nx)];
      h0[j] = (v - net.xu_mean[j]) / net.xu_std[j];
    }
    __syncwarp();
    double *hin = h0, *hout = h1, *Jp = J0, *Jn = J1;
    for (int l = 0; l < net.n_layers; ++l) {
      const int Kin = net.dims[l], N = net.dims[l + 1];
      const bool last = (l == net.n_layers - 1);
      const double *Wt = net.Wt[l];
      for (int j = lane; j < N; j += 32) {
        const double y = net.b[l][j] + dot_col_any(Wt, N, j, hin, Kin);
        hout[j] = last ? y : ampc_act<double>(net.act, y);
        g[j] = last ? 1.0 : ampc_act_grad<double>(net.act, y);
      }
      __syncwarp();
      for (int j = lane; j < N; j += 32) {
        const double gj = g[j];
        if (l == 0) {
          for (int c = 0; c < nin; ++c) Jn[(size_t)j * nin + c] = Wt[(size_t)c * N + j] / net.xu_std[c] * gj;
        } else {
          for (int c0 = 0; c0 < nin; c0 += JAC_CB) {
            double acc0[JAC_CB], acc1[JAC_CB];          // even / odd k, like the two partial sums of the batch routine
#pragma unroll
            for (int q = 0; q < JAC_CB; ++q) { acc0[q] = 0.0; acc1[q] = 0.0; }
            int k = 0;
            for (; k + 2 <= Kin; k += 2) {
              const double w0 = Wt[(size_t)k * N + j], w1 = Wt[(size_t)(k + 1) * N + j];
              const double *r0 = Jp + (size_t)k * nin + c0, *r1 = r0 + nin;
#pragma unroll
              for (int q = 0; q < JAC_CB; ++q)
                if (c0 + q < nin) { acc0[q] = fma(w0, r0[q], acc0[q]); acc1[q] = fma(w1, r1[q], acc1[q]); }
            }
            if (k < Kin) {
              const double w0 = Wt[(size_t)k * N + j];
              const double *r0 = Jp + (size_t)k * nin + c0;
#pragma unroll
              for (int q = 0; q < JAC_CB; ++q)
                if (c0 + q < nin) acc0[q] = fma(w0, r0[q], acc0[q]);
            }
#pragma unroll
            for (int q = 0; q < JAC_CB; ++q)
              if (c0 + q < nin) Jn[(size_t)j * nin + c0 + q] = (acc0[q] + acc1[q]) * gj;
          }
        }
      }
      __syncwarp();
      double *t2 = hin; hin = hout; hout = t2;
      double *t3 = Jp; Jp = Jn; Jn = t3;
    }
    for (int r = lane; r < nx * nin; r += 32) {
      const int a = r / nin, c = r - a * nin;
      Jacs[(size_t)s * nx * nin + r] = Jp[r] * net.dy_std[a] + ((c == a) ? 1.0 : 0.0);
    }
    __syncwarp();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NT) ilqr_kernel(const IlqrParams P) {
  extern __shared__ double sm[];
  const long long t_entry = clock64();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nx = P.nx, nu = P.nu, n = nx + nu, H = P.H, LS = P.ls_max_iter, mw = P.net.max_width;
  // line-search groups: G warps per alpha (2 when they fit), group j = warps [j*G, (j+1)*G)
  const int G = (2 * LS <= NWARPS) ? 2 : 1;
  const int gthr = 32 * G;
  const int grp = warp / G, gl = tid - grp * gthr;
  // shared-memory carve
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
  double *s_h = s_obj + LS + 4;      // LS * 2 * mw: per-group activation ping-pong
  double *s_next = s_h + (size_t)LS * 2 * mw;
  __shared__ int s_flag[4];          // [0]=line search failed, [1]=used idx, [2]=refresh jac
  __shared__ int s_piv[MAX_NU];
  __shared__ double s_lin, s_quad;

  // ---- the network: weights staged into shared memory when they fit (plain loads serve both placements)
  AmpcMlpF64 net = P.net;
  if (P.w_smem) {
    double *s_net = s_next;
    s_next += P.net_doubles;
    for (size_t t = tid; t < P.net_doubles; t += NT) s_net[t] = P.net_blob[t];
    for (int l = 0; l < net.n_layers; ++l) {
      net.Wt[l] = s_net + (P.net.Wt[l] - P.net_blob);
      net.b[l] = s_net + (P.net.b[l] - P.net_blob);
    }
    net.xu_mean = s_net + (P.net.xu_mean - P.net_blob); net.xu_std = s_net + (P.net.xu_std - P.net_blob);
    net.dy_mean = s_net + (P.net.dy_mean - P.net_blob); net.dy_std = s_net + (P.net.dy_std - P.net_blob);
  }
  // ---- trajectories, gains, Jacobians, line-search rollouts: shared memory when they fit
  double *states = P.states, *ctrls = P.ctrls, *Ks = P.Ks, *ks = P.ks, *Jacs = P.Jacs, *ls_states = P.ls_states,
         *ls_ctrls = P.ls_ctrls, *step_cost = P.step_cost;
  if (P.traj_smem) {
    states = s_next; s_next += (size_t)(H + 1) * nx;
    ctrls = s_next; s_next += (size_t)H * nu;
    Ks = s_next; s_next += (size_t)H * nu * nx;
    ks = s_next; s_next += (size_t)H * nu;
    Jacs = s_next; s_next += (size_t)H * nx * n;
    ls_states = s_next; s_next += (size_t)LS * (H + 1) * nx;
    ls_ctrls = s_next; s_next += (size_t)LS * H * nu;
    step_cost = s_next; s_next += (size_t)(LS + 1) * (H + 1);
  }

  // per-warp scratch of the Jacobian refresh
  const size_t jac_per_warp = 3 * (size_t)mw + 2 * (size_t)mw * n;
  double *jac_wk = P.jac_smem ? s_next + (size_t)warp * jac_per_warp : P.hA + (size_t)warp * jac_per_warp;

  for (int t = tid; t < n * n; t += NT) {
    const int r = t / n, c = t - r * n;
    double v = 0.0;
    if (r < nx && c < nx) v = P.dt * (P.Q[r * nx + c] + P.Q[c * nx + r]);
    else if (r >= nx && c >= nx) v = P.dt * (P.R[(r - nx) * nu + (c - nx)] + P.R[(c - nx) * nu + (r - nx)]);
    s_Ct[t] = v;
  }
  for (int t = tid; t < nx * nx; t += NT) {
    const int r = t / nx, c = t - r * nx;
    s_Fs[t] = P.F[r * nx + c] + P.F[c * nx + r];
  }
  for (int t = tid; t < nx; t += NT) states[t] = P.x0[t];
  for (int t = tid; t < H * nu; t += NT) ctrls[t] = P.uguess ? P.uguess[t] : 0.0;
  for (int t = tid; t < P.max_iter; t += NT) P.alpha_idx[t] = -1;
  __syncthreads();

  // ---- initial rollout (ilqr.py:141-147) by line-search group 0; Jacobians are evaluated in one batch afterwards
  if (grp == 0) {
    double *h0 = s_h, *h1 = s_h + mw;
    for (int i = 0; i < H; ++i) {
      for (int j = gl; j < n; j += gthr) {
        const double v = j < nx ? states[(size_t)i * nx + j] : ctrls[(size_t)i * nu + (j - nx)];
        h0[j] = (v - net.xu_mean[j]) / net.xu_std[j];
      }
      group_sync(1, gthr);
      const double *out = group_forward(net, h0, h1, gl, gthr, 1);
      for (int j = gl; j < nx; j += gthr)
        states[(size_t)(i + 1) * nx + j] = states[(size_t)i * nx + j] + (out[j] * net.dy_std[j] + net.dy_mean[j]);
      group_sync(1, gthr);
    }
  }
  __syncthreads();
  jac_batch(P, net, states, ctrls, Jacs, jac_wk, warp, lane);
  step_costs(P, states, ctrls, step_cost, tid, NT);
  __syncthreads();
  double obj = 0.0;       // every thread tracks the same scalars (uniform control flow)
  for (int i = 0; i <= H; ++i) obj += step_cost[i];     // sequential like eval_obj, ilqr.py:124-129
  __syncthreads();

  long long t_mark = clock64();
  const long long t_begin = t_entry;
  unsigned long long cyc[6] = {0, 0, 0, 0, 0, 0};
  auto lap = [&](int slot) { const long long now = clock64(); cyc[slot] += (unsigned long long)(now - t_mark); t_mark = now; };
  cyc[0] = (unsigned long long)(t_mark - t_begin);
  int converged = 0, n_iter = 0, ls_fail = 0;
  for (int itr = 0; itr < P.max_iter; ++itr) {
    n_iter = itr + 1;
    // ---- backward pass (ilqr.py:159-187): ONE warp, warp barriers only (the matrices are (nx+nu)^2)
    if (warp == 0) {
      double *Vn = s_V0, *Vnn = s_V1, *vn = s_v0, *vnn = s_v1;
      for (int t = lane; t < nx * nx; t += 32) Vn[t] = s_Fs[t];
      for (int a = lane; a < nx; a += 32) {
        double acc = 0.0;
        for (int b = 0; b < nx; ++b) acc += s_Fs[a * nx + b] * states[(size_t)H * nx + b];   // no goal: cost.py:208
        vn[a] = acc;
      }
      double lin_acc = 0.0, quad_acc = 0.0;               // lane 0's running sums (ilqr.py:178-179)
      __syncwarp();
      for (int t = H; t >= 1; --t) {
        const double *J = Jacs + (size_t)(t - 1) * nx * n;
        const double *xt = states + (size_t)(t - 1) * nx, *ut = ctrls + (size_t)(t - 1) * nu;
        for (int e = lane; e < nx * n; e += 32) {         // T = Vn @ J
          const int a = e / n, c = e - a * n;
          double acc = 0.0;
          for (int b = 0; b < nx; ++b) acc += Vn[a * nx + b] * J[b * n + c];
          s_T[e] = acc;
        }
        __syncwarp();
        for (int e = lane; e < n * n + n; e += 32) {      // Qt = Ct + J^T T ; qt = ct + J^T vn
          if (e < n * n) {
            const int r = e / n, c = e - r * n;
            double acc = 0.0;
            for (int a = 0; a < nx; ++a) acc += J[a * n + r] * s_T[a * n + c];
            s_Qt[e] = s_Ct[e] + acc;
          } else {
            const int r = e - n * n;
            double ct = 0.0;
            if (r < nx) { for (int b = 0; b < nx; ++b) ct += s_Ct[r * n + b] * (xt[b] - P.goal[b]); }
            else { for (int b = 0; b < nu; ++b) ct += s_Ct[r * n + nx + b] * ut[b]; }
            double acc = 0.0;
            for (int a = 0; a < nx; ++a) acc += J[a * n + r] * vn[a];
            s_qt[r] = ct + acc;
          }
        }
        __syncwarp();
        if (lane == 0) {                                  // LU of Quu with partial pivoting (LAPACK gesv)
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
        for (int col = lane; col <= nx; col += 32) {      // K = -Quu^-1 Qux, k = -Quu^-1 qu: one right-hand side per lane
          double y[MAX_NU];
          for (int r = 0; r < nu; ++r) y[r] = (col < nx) ? s_Qt[(nx + r) * n + col] : s_qt[nx + r];
          for (int c = 0; c < nu; ++c) { const int pc = s_piv[c]; if (pc != c) { double tmp = y[c]; y[c] = y[pc]; y[pc] = tmp; } }
          for (int r = 1; r < nu; ++r) for (int q = 0; q < r; ++q) y[r] -= s_LU[r * nu + q] * y[q];
          for (int r = nu - 1; r >= 0; --r) { for (int q = r + 1; q < nu; ++q) y[r] -= s_LU[r * nu + q] * y[q]; y[r] /= s_LU[r * nu + r]; }
          for (int r = 0; r < nu; ++r) { if (col < nx) s_K[r * nx + col] = -y[r]; else s_k[r] = -y[r]; }
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
        for (int e = lane; e < nx * nx + nx; e += 32) {   // value update (ilqr.py:186-187)
          if (e < nx * nx) {
            const int a = e / nx, b = e - a * nx;
            double acc = s_Qt[a * n + b];
            for (int r = 0; r < nu; ++r) acc += s_Qt[a * n + nx + r] * s_K[r * nx + b];
            for (int r = 0; r < nu; ++r) acc += s_K[r * nx + a] * s_Qt[(nx + r) * n + b];
            for (int r = 0; r < nu; ++r) {
              double row = 0.0;
              for (int c = 0; c < nu; ++c) row += s_Qt[(nx + r) * n + nx + c] * s_K[c * nx + b];
              acc += s_K[r * nx + a] * row;
            }
            Vnn[e] = acc;
          } else {
            const int a = e - nx * nx;
            double acc = s_qt[a];
            for (int r = 0; r < nu; ++r) acc += s_Qt[a * n + nx + r] * s_k[r];
            for (int r = 0; r < nu; ++r) {
              double inner = s_qt[nx + r];
              for (int c = 0; c < nu; ++c) inner += s_Qt[(nx + r) * n + nx + c] * s_k[c];
              acc += s_K[r * nx + a] * inner;
            }
            vnn[a] = acc;
          }
        }
        __syncwarp();
        double *tp = Vn; Vn = Vnn; Vnn = tp;
        tp = vn; vn = vnn; vnn = tp;
      }
      if (lane == 0) { s_lin = lin_acc; s_quad = quad_acc; }
    }
    __syncthreads();
    lap(1);
    const double lin_cost_reduce = s_lin, quad_cost_reduce = s_quad;
    double ksq = 0.0;
    for (int t = tid; t < H * nu; t += NT) ksq += ks[t] * ks[t];
    const double ks_norm = sqrt(block_sum(ksq, s_red, tid));

    // ---- line-search rollouts (ilqr.py:190-205): alpha j is rolled out by group j on its own, H sequential steps with
    //      group barriers only; alphas beyond the number of groups are taken in further rounds
    const int n_groups = NWARPS / G;
    for (int j = grp; j < LS; j += n_groups) {
      const int bar = 1 + (grp % 15);
      double *h0 = s_h + (size_t)(grp % LS) * 2 * mw, *h1 = h0 + mw;
      double *xs = ls_states + (size_t)j * (H + 1) * nx, *us = ls_ctrls + (size_t)j * H * nu;
      const double alpha = P.alphas[j];
      for (int a = gl; a < nx; a += gthr) xs[a] = P.x0[a];
      group_sync(bar, gthr);
      for (int i = 0; i < H; ++i) {
        const double *xi = xs + (size_t)i * nx;
        for (int a = gl; a < nu; a += gthr) {
          double fb = 0.0;
          for (int b = 0; b < nx; ++b) fb += Ks[((size_t)i * nu + a) * nx + b] * (xi[b] - states[(size_t)i * nx + b]);
          double u = alpha * ks[(size_t)i * nu + a] + ctrls[(size_t)i * nu + a] + fb;
          if (P.bounded) u = fmin(fmax(u, P.umin[a]), P.umax[a]);     // np.clip, ilqr.py:203-204
          us[(size_t)i * nu + a] = u;
        }
        group_sync(bar, gthr);
        for (int c = gl; c < n; c += gthr) {
          const double v = c < nx ? xi[c] : us[(size_t)i * nu + (c - nx)];
          h0[c] = (v - net.xu_mean[c]) / net.xu_std[c];
        }
        group_sync(bar, gthr);
        const double *out = group_forward(net, h0, h1, gl, gthr, bar);
        for (int a = gl; a < nx; a += gthr)
          xs[(size_t)(i + 1) * nx + a] = xi[a] + (out[a] * net.dy_std[a] + net.dy_mean[a]);
        group_sync(bar, gthr);
      }
    }
    __syncthreads();
    lap(2);
    // objective of every alpha: per-step costs in parallel, then a sequential sum per alpha
    for (int t = tid; t < LS * (H + 1); t += NT) {
      const int j = t / (H + 1), i = t - j * (H + 1);
      const double *xs = ls_states + (size_t)j * (H + 1) * nx, *us = ls_ctrls + (size_t)j * H * nu;
      double c;
      if (i < H) c = P.dt * (quad_form(P.Q, xs + (size_t)i * nx, P.goal, nx) + quad_form(P.R, us + (size_t)i * nu, nullptr, nu));
      else c = quad_form(P.F, xs + (size_t)H * nx, P.goalF, nx);
      step_cost[(H + 1) + t] = c;
    }
    __syncthreads();
    if (tid < LS) {
      double o = 0.0;
      for (int i = 0; i <= H; ++i) o += step_cost[(H + 1) + tid * (H + 1) + i];
      s_obj[tid] = o;
    }
    __syncthreads();
    // ---- backtracking acceptance (ilqr.py:208-238), thread 0 decides
    if (tid == 0) {
      int best_idx = -1, used = -1, have_best = 0;
      double best_obj = INFINITY, new_obj = 0.0;
      for (int l = 0; l < LS; ++l) {
        used = l;
        new_obj = s_obj[l];
        const double alpha = P.alphas[l];
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
    if (s_flag[2]) jac_batch(P, net, nxs, nus, Jacs, jac_wk, warp, lane);   // ilqr.py:232 (stale Jacobians otherwise, as in the reference)
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
  if (P.traj_smem) {                                     // outputs back to global memory
    __syncthreads();
    for (int t = tid; t < (H + 1) * nx; t += NT) P.states[t] = states[t];
    for (int t = tid; t < H * nu; t += NT) { P.ctrls[t] = ctrls[t]; P.ks[t] = ks[t]; }
    for (int t = tid; t < H * nu * nx; t += NT) P.Ks[t] = Ks[t];
  }
  if (tid == 0) {
    P.info[0] = converged; P.info[1] = n_iter; P.info[2] = ls_fail;
    lap(5);
    for (int q = 0; q < 6; ++q) P.prof[q] = cyc[q];
    P.prof[6] = (unsigned long long)(clock64() - t_begin);
    P.prof[7] = (unsigned long long)n_iter;
  }
}

}  // namespace

struct ampc_ilqr {
  ampc_ilqr_cfg cfg;
  IlqrParams P;
  int device = 0;
  double *d_blob = nullptr;   // MLP weights
  double *d_work = nullptr;   // everything else
  int *d_int = nullptr;
  unsigned long long *d_prof = nullptr;
  size_t smem = 0;
  size_t o_x0 = 0, o_ug = 0;
};

extern "C" int ampc_ilqr_create(ampc_ilqr **out, const ampc_ilqr_cfg *cfg, const ampc_mlp_desc *mlp,
                                const ampc_quad_cost *cost) {
  AMPC_REQUIRE(out && cfg && mlp && cost, AMPC_ERR_INVALID, "null argument");
  *out = nullptr;
  AMPC_REQUIRE(cfg->H >= 1 && cfg->nx >= 1 && cfg->nu >= 1 && cfg->nu <= MAX_NU, AMPC_ERR_INVALID,
               "bad iLQR dims H=%d nx=%d nu=%d (nu <= %d)", cfg->H, cfg->nx, cfg->nu, MAX_NU);
  AMPC_REQUIRE(cfg->max_iter >= 1 && cfg->ls_max_iter >= 1 && cfg->ls_max_iter <= MAX_LS, AMPC_ERR_INVALID,
               "bad iteration limits (1 <= ls_max_iter <= %d)", MAX_LS);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  AMPC_REQUIRE(ce == cudaSuccess && ndev > 0, AMPC_ERR_CUDA, "no CUDA device: libampc_b200 has no CPU fallback (%s)",
               cudaGetErrorString(ce));
  AMPC_REQUIRE(cfg->device >= 0 && cfg->device < ndev, AMPC_ERR_INVALID, "device %d of %d", cfg->device, ndev);
  AMPC_CUDA_CHECK(cudaSetDevice(cfg->device));
  ampc_ilqr *h = new ampc_ilqr();
  h->cfg = *cfg;
  h->device = cfg->device;
  IlqrParams &P = h->P;
  memset(&P, 0, sizeof(P));
  int rc = ampc_mlp_f64_upload(mlp, cfg->nx, cfg->nu, &P.net, &h->d_blob);
  if (rc) { delete h; return rc; }
  const int H = cfg->H, nx = cfg->nx, nu = cfg->nu, n = nx + nu, LS = cfg->ls_max_iter, mw = P.net.max_width;
  P.H = H; P.nx = nx; P.nu = nu; P.bounded = cfg->bounded; P.max_iter = cfg->max_iter; P.ls_max_iter = LS;
  P.dt = cfg->dt; P.ls_discount = cfg->ls_discount; P.ls_cost_threshold = cfg->ls_cost_threshold;
  P.u_threshold = cfg->u_threshold;
  size_t off = 0;
  auto take = [&](size_t cnt) { size_t o = off; off += (cnt + 1) & ~(size_t)1; return o; };
  const size_t oQ = take(nx * nx), oR = take(nu * nu), oF = take(nx * nx), og = take(nx), ogF = take(nx), oumin = take(nu), oumax = take(nu), oal = take(LS);
  h->o_x0 = take(nx); h->o_ug = take((size_t)H * nu);
  const size_t ost = take((size_t)(H + 1) * nx), oct = take((size_t)H * nu), oKs = take((size_t)H * nu * nx), oks = take((size_t)H * nu);
  const size_t oJ = take((size_t)H * nx * n), ols = take((size_t)LS * (H + 1) * nx), olc = take((size_t)LS * H * nu);
  const size_t osc = take((size_t)(LS + 1) * (H + 1));
  const size_t jac_per_warp = 3 * (size_t)mw + 2 * (size_t)mw * n;
  const size_t ohA = take((size_t)NWARPS * jac_per_warp);
  cudaError_t e = cudaMalloc(&h->d_work, off * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(h->d_work, 0, off * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_int, (3 + cfg->max_iter) * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_prof, 8 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(h->d_prof, 0, 8 * sizeof(unsigned long long));
  std::vector<double> hc(h->o_x0, 0.0);
  for (int i = 0; i < nx * nx; ++i) { hc[oQ + i] = cost->Q[i]; hc[oF + i] = cost->F[i]; }
  for (int i = 0; i < nu * nu; ++i) hc[oR + i] = cost->R[i];
  for (int i = 0; i < nx; ++i) { hc[og + i] = cost->goal[i]; hc[ogF + i] = cost->goal_term ? cost->goal_term[i] : cost->goal[i]; }
  for (int i = 0; i < nu; ++i) { hc[oumin + i] = cost->umin[i]; hc[oumax + i] = cost->umax[i]; }
  for (int i = 0; i < LS; ++i) hc[oal + i] = pow(cfg->ls_discount, (double)i);   // ls_discount**i, ilqr.py:196
  if (e == cudaSuccess) e = cudaMemcpy(h->d_work, hc.data(), hc.size() * sizeof(double), cudaMemcpyHostToDevice);
  double *w = h->d_work;
  P.Q = w + oQ; P.R = w + oR; P.F = w + oF; P.goal = w + og; P.goalF = w + ogF; P.umin = w + oumin; P.umax = w + oumax; P.alphas = w + oal;
  P.x0 = w + h->o_x0; P.uguess = nullptr;
  P.states = w + ost; P.ctrls = w + oct; P.Ks = w + oKs; P.ks = w + oks; P.Jacs = w + oJ;
  P.ls_states = w + ols; P.ls_ctrls = w + olc; P.step_cost = w + osc;
  P.hA = w + ohA;
  P.info = h->d_int; P.alpha_idx = h->d_int + 3; P.prof = h->d_prof;
  {
    // shared memory: the small matrices always; the network and the trajectories / gains / Jacobians / line-search
    // rollouts when they fit next to them (the cartpole problem: 37 KB + 37 KB)
    size_t fixed = (size_t)n * n * 2 + (size_t)nx * nx * 3 + 2 * nx + (size_t)nx * n + n + (size_t)nu * nx + nu +
                   (size_t)nu * nu + NWARPS + LS + 4 + (size_t)LS * 2 * mw;
    size_t netd = 2 * (size_t)(nx + nu) + 2 * (size_t)nx;
    for (int l = 0; l < mlp->n_layers; ++l) netd += (size_t)mlp->dims[l] * mlp->dims[l + 1] + mlp->dims[l + 1];
    const size_t traj = (size_t)(H + 1) * nx + (size_t)H * nu * 2 + (size_t)H * nu * nx + (size_t)H * nx * n +
                        (size_t)LS * (H + 1) * nx + (size_t)LS * H * nu + (size_t)(LS + 1) * (H + 1);
    int max_optin = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
    const size_t cap = (size_t)max_optin > 2048 ? ((size_t)max_optin - 2048) / sizeof(double) : 0;
    P.net_doubles = netd;
    P.net_blob = h->d_blob;
    P.w_smem = (fixed + netd <= cap) ? 1 : 0;
    if (P.w_smem) fixed += netd;
    P.traj_smem = (fixed + traj <= cap) ? 1 : 0;
    if (P.traj_smem) fixed += traj;
    const size_t jacs = (size_t)NWARPS * jac_per_warp;
    P.jac_smem = (fixed + jacs <= cap) ? 1 : 0;
    if (P.jac_smem) fixed += jacs;
    if (getenv("AMPC_ILQR_NO_SMEM")) {   // debugging / A-B: everything in global memory like the round-1 kernel
      fixed -= (P.w_smem ? netd : 0) + (P.traj_smem ? traj : 0) + (P.jac_smem ? jacs : 0);
      P.w_smem = P.traj_smem = P.jac_smem = 0;
    }
    h->smem = fixed * sizeof(double);
    AMPC_REQUIRE(fixed <= cap || e != cudaSuccess, AMPC_ERR_UNSUPPORTED, "iLQR: %zu B of shared memory needed", h->smem);
  }
  if (e == cudaSuccess) e = ampc_raise_smem_limit((const void *)ilqr_kernel, h->smem);
  if (e != cudaSuccess) {
    ampc_set_error("iLQR create: %s", cudaGetErrorString(e));
    cudaFree(h->d_blob); cudaFree(h->d_work); cudaFree(h->d_int); cudaFree(h->d_prof);
    delete h;
    return AMPC_ERR_CUDA;
  }
  *out = h;
  return AMPC_OK;
}

extern "C" int ampc_ilqr_destroy(ampc_ilqr *h) {
  if (!h) return AMPC_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_blob); cudaFree(h->d_work); cudaFree(h->d_int); cudaFree(h->d_prof);
  delete h;
  return AMPC_OK;
}

// Debug tap, no reference counterpart: SM cycles thread 0 of the last solve spent per phase (see IlqrParams::prof).
extern "C" int ampc_ilqr_debug_profile(ampc_ilqr *h, unsigned long long *out8) {
  AMPC_REQUIRE(h && out8, AMPC_ERR_INVALID, "null argument");
  AMPC_CUDA_CHECK(cudaSetDevice(h->device));
  AMPC_CUDA_CHECK(cudaDeviceSynchronize());
  AMPC_CUDA_CHECK(cudaMemcpy(out8, h->d_prof, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return AMPC_OK;
}

// Re-runs the solve on whatever x0 / uguess the last solve_host left on the device, asynchronously on `stream`: no
// host copies.  Used to time the kernel alone (bench.py --workload c4); outputs stay on the device.
extern "C" int ampc_ilqr_launch(ampc_ilqr *h, void *stream) {
  AMPC_REQUIRE(h, AMPC_ERR_INVALID, "null handle");
  AMPC_CUDA_CHECK(cudaSetDevice(h->device));
  ilqr_kernel<<<1, NT, h->smem, (cudaStream_t)stream>>>(h->P);
  ampc_count_launch();
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
  ilqr_kernel<<<1, NT, h->smem>>>(P);
  ampc_count_launch();
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
