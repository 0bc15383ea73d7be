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

constexpr int NT = 512;
constexpr int MAX_NU = 16;

struct IlqrParams {
  AmpcMlpF64 net;
  int H, nx, nu, bounded, max_iter, ls_max_iter;
  double dt, ls_discount, ls_cost_threshold, u_threshold;
  const double *Q, *R, *F, *goal, *goalF, *umin, *umax, *alphas;   // device (goalF: goal of the terminal term)
  const double *x0, *uguess;                       // device (uguess may be null)
  // scratch / outputs (device)
  double *states, *ctrls, *Ks, *ks, *Jacs, *ls_states, *ls_ctrls, *step_cost;
  double *hA, *hB, *hG, *JA, *JB;
  int *info, *alpha_idx;
};

__device__ __forceinline__ double block_sum(double v, double *s_red, int tid) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < NT / 32; ++w) r += s_red[w];
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

// Jacobians of x' = x + dy(x,u) at (xs[i], us[i]) for i < H  ->  Jacs (H, nx, nx+nu)   (mlp.py:281-305)
__device__ void jac_batch(const IlqrParams &P, const double *xs, const double *us, int tid) {
  const AmpcMlpF64 &net = P.net;
  const int nx = P.nx, nu = P.nu, nin = nx + nu, H = P.H, mw = net.max_width;
  for (int t = tid; t < H * nin; t += NT) {
    const int s = t / nin, j = t - s * nin;
    const double v = j < nx ? xs[(size_t)s * nx + j] : us[(size_t)s * nu + (j - nx)];
    P.hA[(size_t)s * mw + j] = (v - __ldg(net.xu_mean + j)) / __ldg(net.xu_std + j);
  }
  __syncthreads();
  const double *J;
  ampc_mlp_f64_forward_jac_batch(net, H, P.hA, P.hB, P.hG, mw, P.JA, P.JB, mw * nin, &J, tid, NT);
  for (int t = tid; t < H * nx * nin; t += NT) {
    const int s = t / (nx * nin), r = t - s * (nx * nin);
    const int a = r / nin, c = r - a * nin;
    P.Jacs[t] = J[(size_t)s * mw * nin + r] * __ldg(net.dy_std + a) + ((c == a) ? 1.0 : 0.0);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NT) ilqr_kernel(const IlqrParams P) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x;
  const int nx = P.nx, nu = P.nu, n = nx + nu, H = P.H, LS = P.ls_max_iter, mw = P.net.max_width;
  const AmpcMlpF64 &net = P.net;
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
  double *s_red = s_k + nu;          // NT/32
  double *s_obj = s_red + NT / 32;   // LS + 4
  __shared__ int s_flag[4];          // [0]=continue loop, [1]=used idx, [2]=refresh jac
  __shared__ double s_lin, s_quad;

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
  for (int t = tid; t < nx; t += NT) P.states[t] = P.x0[t];
  for (int t = tid; t < H * nu; t += NT) P.ctrls[t] = P.uguess ? P.uguess[t] : 0.0;
  for (int t = tid; t < P.max_iter; t += NT) P.alpha_idx[t] = -1;
  __syncthreads();

  // ---- initial rollout (ilqr.py:141-147); Jacobians are evaluated in one batch afterwards
  for (int i = 0; i < H; ++i) {
    for (int j = tid; j < n; j += NT) {
      const double v = j < nx ? P.states[(size_t)i * nx + j] : P.ctrls[(size_t)i * nu + (j - nx)];
      P.hA[j] = (v - __ldg(net.xu_mean + j)) / __ldg(net.xu_std + j);
    }
    __syncthreads();
    const double *out = ampc_mlp_f64_forward_batch(net, 1, P.hA, P.hB, mw, tid, NT);
    for (int j = tid; j < nx; j += NT)
      P.states[(size_t)(i + 1) * nx + j] = P.states[(size_t)i * nx + j] + (out[j] * __ldg(net.dy_std + j) + __ldg(net.dy_mean + j));
    __syncthreads();
  }
  jac_batch(P, P.states, P.ctrls, tid);
  step_costs(P, P.states, P.ctrls, P.step_cost, tid, NT);
  __syncthreads();
  double obj = 0.0;       // every thread tracks the same scalars (uniform control flow)
  for (int i = 0; i <= H; ++i) obj += P.step_cost[i];     // sequential like eval_obj, ilqr.py:124-129
  __syncthreads();

  int converged = 0, n_iter = 0, ls_fail = 0;
  for (int itr = 0; itr < P.max_iter; ++itr) {
    n_iter = itr + 1;
    // ---- backward pass (ilqr.py:159-187)
    double *Vn = s_V0, *Vnn = s_V1, *vn = s_v0, *vnn = s_v1;
    for (int t = tid; t < nx * nx; t += NT) Vn[t] = s_Fs[t];
    for (int a = tid; a < nx; a += NT) {
      double acc = 0.0;
      for (int b = 0; b < nx; ++b) acc += s_Fs[a * nx + b] * P.states[(size_t)H * nx + b];   // no goal: cost.py:208
      vn[a] = acc;
    }
    if (tid == 0) { s_lin = 0.0; s_quad = 0.0; }
    __syncthreads();
    for (int t = H; t >= 1; --t) {
      const double *J = P.Jacs + (size_t)(t - 1) * nx * n;
      const double *xt = P.states + (size_t)(t - 1) * nx, *ut = P.ctrls + (size_t)(t - 1) * nu;
      for (int e = tid; e < nx * n; e += NT) {         // T = Vn @ J
        const int a = e / n, c = e - a * n;
        double acc = 0.0;
        for (int b = 0; b < nx; ++b) acc += Vn[a * nx + b] * J[b * n + c];
        s_T[e] = acc;
      }
      __syncthreads();
      for (int e = tid; e < n * n + n; e += NT) {      // Qt = Ct + J^T T ; qt = ct + J^T vn
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
      __syncthreads();
      if (tid == 0) {                                   // K = -Quu^-1 Qux, k = -Quu^-1 qu   (ilqr.py:176-177)
        double A[MAX_NU][MAX_NU];
        int piv[MAX_NU];
        for (int r = 0; r < nu; ++r) for (int c = 0; c < nu; ++c) A[r][c] = s_Qt[(nx + r) * n + nx + c];
        for (int c = 0; c < nu; ++c) {                  // LU with partial pivoting (LAPACK gesv)
          int pr = c; double best = fabs(A[c][c]);
          for (int r = c + 1; r < nu; ++r) if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); pr = r; }
          piv[c] = pr;
          if (pr != c) for (int q = 0; q < nu; ++q) { double tmp = A[c][q]; A[c][q] = A[pr][q]; A[pr][q] = tmp; }
          for (int r = c + 1; r < nu; ++r) {
            A[r][c] /= A[c][c];
            for (int q = c + 1; q < nu; ++q) A[r][q] -= A[r][c] * A[c][q];
          }
        }
        for (int col = 0; col <= nx; ++col) {
          double y[MAX_NU];
          for (int r = 0; r < nu; ++r) y[r] = (col < nx) ? s_Qt[(nx + r) * n + col] : s_qt[nx + r];
          for (int c = 0; c < nu; ++c) { if (piv[c] != c) { double tmp = y[c]; y[c] = y[piv[c]]; y[piv[c]] = tmp; } }
          for (int r = 1; r < nu; ++r) for (int q = 0; q < r; ++q) y[r] -= A[r][q] * y[q];
          for (int r = nu - 1; r >= 0; --r) { for (int q = r + 1; q < nu; ++q) y[r] -= A[r][q] * y[q]; y[r] /= A[r][r]; }
          for (int r = 0; r < nu; ++r) { if (col < nx) s_K[r * nx + col] = -y[r]; else s_k[r] = -y[r]; }
        }
        double lin = 0.0, quad = 0.0;
        for (int r = 0; r < nu; ++r) {
          lin += s_qt[nx + r] * s_k[r];
          double row = 0.0;
          for (int c = 0; c < nu; ++c) row += s_Qt[(nx + r) * n + nx + c] * s_k[c];
          quad += s_k[r] * row;
        }
        s_lin += lin; s_quad += quad;                   // ilqr.py:178-179
      }
      __syncthreads();
      for (int e = tid; e < nu * nx + nu; e += NT) {
        if (e < nu * nx) P.Ks[(size_t)(t - 1) * nu * nx + e] = s_K[e];
        else P.ks[(size_t)(t - 1) * nu + (e - nu * nx)] = s_k[e - nu * nx];
      }
      for (int e = tid; e < nx * nx + nx; e += NT) {    // value update (ilqr.py:186-187)
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
      __syncthreads();
      double *tp = Vn; Vn = Vnn; Vnn = tp;
      tp = vn; vn = vnn; vnn = tp;
    }
    const double lin_cost_reduce = s_lin, quad_cost_reduce = s_quad;
    double ksq = 0.0;
    for (int t = tid; t < H * nu; t += NT) ksq += P.ks[t] * P.ks[t];
    const double ks_norm = sqrt(block_sum(ksq, s_red, tid));

    // ---- line-search rollouts for all alphas (ilqr.py:190-205)
    for (int t = tid; t < LS * nx; t += NT) {
      const int j = t / nx, a = t - j * nx;
      P.ls_states[(size_t)j * (H + 1) * nx + a] = P.x0[a];
    }
    __syncthreads();
    for (int i = 0; i < H; ++i) {
      for (int t = tid; t < LS * nu; t += NT) {
        const int j = t / nu, a = t - j * nu;
        const double alpha = P.alphas[j];
        const double *xs = P.ls_states + ((size_t)j * (H + 1) + i) * nx;
        double fb = 0.0;
        for (int b = 0; b < nx; ++b) fb += P.Ks[((size_t)i * nu + a) * nx + b] * (xs[b] - P.states[(size_t)i * nx + b]);
        double u = alpha * P.ks[(size_t)i * nu + a] + P.ctrls[(size_t)i * nu + a] + fb;
        if (P.bounded) u = fmin(fmax(u, P.umin[a]), P.umax[a]);     // np.clip, ilqr.py:203-204
        P.ls_ctrls[((size_t)j * H + i) * nu + a] = u;
      }
      __syncthreads();
      for (int t = tid; t < LS * n; t += NT) {
        const int j = t / n, c = t - j * n;
        const double v = c < nx ? P.ls_states[((size_t)j * (H + 1) + i) * nx + c] : P.ls_ctrls[((size_t)j * H + i) * nu + (c - nx)];
        P.hA[(size_t)j * mw + c] = (v - __ldg(net.xu_mean + c)) / __ldg(net.xu_std + c);
      }
      __syncthreads();
      const double *out = ampc_mlp_f64_forward_batch(net, LS, P.hA, P.hB, mw, tid, NT);
      for (int t = tid; t < LS * nx; t += NT) {
        const int j = t / nx, a = t - j * nx;
        P.ls_states[((size_t)j * (H + 1) + i + 1) * nx + a] =
            P.ls_states[((size_t)j * (H + 1) + i) * nx + a] + (out[(size_t)j * mw + a] * __ldg(net.dy_std + a) + __ldg(net.dy_mean + a));
      }
      __syncthreads();
    }
    // objective of every alpha: per-step costs in parallel, then a sequential sum per alpha
    for (int t = tid; t < LS * (H + 1); t += NT) {
      const int j = t / (H + 1), i = t - j * (H + 1);
      const double *xs = P.ls_states + (size_t)j * (H + 1) * nx, *us = P.ls_ctrls + (size_t)j * H * nu;
      double c;
      if (i < H) c = P.dt * (quad_form(P.Q, xs + (size_t)i * nx, P.goal, nx) + quad_form(P.R, us + (size_t)i * nu, nullptr, nu));
      else c = quad_form(P.F, xs + (size_t)H * nx, P.goalF, nx);
      P.step_cost[(H + 1) + t] = c;
    }
    __syncthreads();
    if (tid < LS) {
      double o = 0.0;
      for (int i = 0; i <= H; ++i) o += P.step_cost[(H + 1) + tid * (H + 1) + i];
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
    if (s_flag[0]) { ls_fail = 1; break; }
    const int used = s_flag[1];
    const double *nxs = P.ls_states + (size_t)used * (H + 1) * nx, *nus = P.ls_ctrls + (size_t)used * H * nu;
    if (s_flag[2]) jac_batch(P, nxs, nus, tid);          // ilqr.py:232 (stale Jacobians otherwise, as in the reference)
    if (tid == 0) P.alpha_idx[itr] = used;
    double dsq = 0.0;
    for (int t = tid; t < H * nu; t += NT) { const double d = nus[t] - P.ctrls[t]; dsq += d * d; }
    const double du_norm = sqrt(block_sum(dsq, s_red, tid));   // ilqr.py:246
    if (du_norm < P.u_threshold) converged = 1;
    for (int t = tid; t < (H + 1) * nx; t += NT) P.states[t] = nxs[t];
    for (int t = tid; t < H * nu; t += NT) P.ctrls[t] = nus[t];
    obj = s_obj[LS];
    __syncthreads();
    if (converged) break;
  }
  if (tid == 0) { P.info[0] = converged; P.info[1] = n_iter; P.info[2] = ls_fail; }
}

}  // namespace

struct ampc_ilqr {
  ampc_ilqr_cfg cfg;
  IlqrParams P;
  int device = 0;
  double *d_blob = nullptr;   // MLP weights
  double *d_work = nullptr;   // everything else
  int *d_int = nullptr;
  size_t smem = 0;
  size_t o_x0 = 0, o_ug = 0;
};

extern "C" int ampc_ilqr_create(ampc_ilqr **out, const ampc_ilqr_cfg *cfg, const ampc_mlp_desc *mlp,
                                const ampc_quad_cost *cost) {
  AMPC_REQUIRE(out && cfg && mlp && cost, AMPC_ERR_INVALID, "null argument");
  *out = nullptr;
  AMPC_REQUIRE(cfg->H >= 1 && cfg->nx >= 1 && cfg->nu >= 1 && cfg->nu <= MAX_NU, AMPC_ERR_INVALID,
               "bad iLQR dims H=%d nx=%d nu=%d (nu <= %d)", cfg->H, cfg->nx, cfg->nu, MAX_NU);
  AMPC_REQUIRE(cfg->max_iter >= 1 && cfg->ls_max_iter >= 1 && cfg->ls_max_iter <= 32, AMPC_ERR_INVALID,
               "bad iteration limits");
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
  const int nb = H > LS ? H : LS;   // widest MLP batch (Jacobian refresh over H steps / LS line-search rollouts)
  size_t off = 0;
  auto take = [&](size_t cnt) { size_t o = off; off += (cnt + 1) & ~(size_t)1; return o; };
  const size_t oQ = take(nx * nx), oR = take(nu * nu), oF = take(nx * nx), og = take(nx), ogF = take(nx), oumin = take(nu), oumax = take(nu), oal = take(LS);
  h->o_x0 = take(nx); h->o_ug = take((size_t)H * nu);
  const size_t ost = take((size_t)(H + 1) * nx), oct = take((size_t)H * nu), oKs = take((size_t)H * nu * nx), oks = take((size_t)H * nu);
  const size_t oJ = take((size_t)H * nx * n), ols = take((size_t)LS * (H + 1) * nx), olc = take((size_t)LS * H * nu);
  const size_t osc = take((size_t)(LS + 1) * (H + 1));
  const size_t ohA = take((size_t)nb * mw), ohB = take((size_t)nb * mw), ohG = take((size_t)nb * mw);
  const size_t oJA = take((size_t)H * mw * n), oJB = take((size_t)H * mw * n);
  cudaError_t e = cudaMalloc(&h->d_work, off * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(h->d_work, 0, off * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_int, (3 + cfg->max_iter) * sizeof(int));
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
  P.hA = w + ohA; P.hB = w + ohB; P.hG = w + ohG; P.JA = w + oJA; P.JB = w + oJB;
  P.info = h->d_int; P.alpha_idx = h->d_int + 3;
  h->smem = ((size_t)n * n * 2 + (size_t)nx * nx * 3 + 2 * nx + (size_t)nx * n + n + (size_t)nu * nx + nu + NT / 32 + LS + 4) * sizeof(double);
  if (e == cudaSuccess) e = ampc_raise_smem_limit((const void *)ilqr_kernel, h->smem);
  if (e != cudaSuccess) {
    ampc_set_error("iLQR create: %s", cudaGetErrorString(e));
    cudaFree(h->d_blob); cudaFree(h->d_work); cudaFree(h->d_int);
    delete h;
    return AMPC_ERR_CUDA;
  }
  *out = h;
  return AMPC_OK;
}

extern "C" int ampc_ilqr_destroy(ampc_ilqr *h) {
  if (!h) return AMPC_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_blob); cudaFree(h->d_work); cudaFree(h->d_int);
  delete h;
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
