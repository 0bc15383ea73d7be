// float64 MLP device routines shared by mlp_ops.cu and ilqr.cu.
// Semantics: autompc/sysid/mlp.py:55-59 (ForwardNet.forward) and the closed form
// of the Jacobian back-propagation in mlp.py:238-305.
#pragma once
#include "ampc_common.cuh"

struct AmpcMlpF64 {
  int n_layers, act, nx, nu, max_width;
  int dims[AMPC_MAX_LAYERS + 1];
  const double *Wt[AMPC_MAX_LAYERS];  // [in][out] (transposed torch.nn.Linear.weight)
  const double *b[AMPC_MAX_LAYERS];
  const double *xu_mean, *xu_std, *dy_mean, *dy_std;
};

int ampc_mlp_f64_upload(const ampc_mlp_desc *mlp, int nx, int nu, AmpcMlpF64 *net, double **blob_out);

// dot(Wt[:, j], h) with four independent partial sums (hides fp64 FMA latency)
__device__ __forceinline__ double ampc_dot_col(const double *__restrict__ Wt, int N, int j, const double *h, int Kin) {
  double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
  int k = 0;
  for (; k + 4 <= Kin; k += 4) {
    p0 = fma(__ldg(Wt + (size_t)(k + 0) * N + j), h[k + 0], p0);
    p1 = fma(__ldg(Wt + (size_t)(k + 1) * N + j), h[k + 1], p1);
    p2 = fma(__ldg(Wt + (size_t)(k + 2) * N + j), h[k + 2], p2);
    p3 = fma(__ldg(Wt + (size_t)(k + 3) * N + j), h[k + 3], p3);
  }
  for (; k < Kin; ++k) p0 = fma(__ldg(Wt + (size_t)k * N + j), h[k], p0);
  return (p0 + p1) + (p2 + p3);
}

// Batched forward.  `hA` holds B z-scored inputs, sample s at hA + s*stride; hA/hB ping-pong.
// Must be called by all `nthr` threads of the CTA (contains __syncthreads).  Returns the
// buffer base holding the raw network outputs (sample s at ret + s*stride).
__device__ __forceinline__ const double *ampc_mlp_f64_forward_batch(const AmpcMlpF64 &net, int B, double *hA,
                                                                    double *hB, int stride, int tid, int nthr) {
  double *hin = hA, *hout = hB;
  for (int l = 0; l < net.n_layers; ++l) {
    const int Kin = net.dims[l], N = net.dims[l + 1];
    const bool last = (l == net.n_layers - 1);
    for (int t = tid; t < B * N; t += nthr) {
      const int s = t / N, j = t - s * N;
      const double y = __ldg(net.b[l] + j) + ampc_dot_col(net.Wt[l], N, j, hin + (size_t)s * stride, Kin);
      hout[(size_t)s * stride + j] = last ? y : ampc_act<double>(net.act, y);
    }
    __syncthreads();
    double *t2 = hin; hin = hout; hout = t2;
  }
  return hin;
}

// Sample-blocked forward for the batch kernels (pred_batch / k-step rollouts at scale): a CTA owns B <= 8 samples and a
// thread computes one output neuron for FOUR of them at a time, so every weight read (L2 / L1; the 3x256 network is
// 1.1 MB of float64) feeds four multiply-adds from registers and eight per CTA instead of one.  Per sample the
// arithmetic is ampc_dot_col's, operation for operation (four partial sums over k mod 4, remainder into the first,
// (p0 + p1) + (p2 + p3), bias added last), so the result is bit-identical to the one-sample routine above.
// Samples beyond B must hold finite inputs (zeros); their outputs are computed and ignored by the caller.
constexpr int AMPC_MLP_SB = 8;   // samples per CTA
__device__ __forceinline__ const double *ampc_mlp_f64_forward_blocked(const AmpcMlpF64 &net, double *hA, double *hB,
                                                                      int stride, int tid, int nthr) {
  constexpr int SPT = 4, G = AMPC_MLP_SB / SPT;
  double *hin = hA, *hout = hB;
  for (int l = 0; l < net.n_layers; ++l) {
    const int Kin = net.dims[l], N = net.dims[l + 1];
    const bool last = (l == net.n_layers - 1);
    const double *__restrict__ Wt = net.Wt[l];
    for (int t = tid; t < G * N; t += nthr) {
      const int g = t / N, j = t - g * N;
      const double *h = hin + (size_t)(g * SPT) * stride;
      double p[SPT][4];
#pragma unroll
      for (int s = 0; s < SPT; ++s) p[s][0] = p[s][1] = p[s][2] = p[s][3] = 0.0;
      int k = 0;
      for (; k + 4 <= Kin; k += 4) {
        const double w0 = __ldg(Wt + (size_t)(k + 0) * N + j), w1 = __ldg(Wt + (size_t)(k + 1) * N + j);
        const double w2 = __ldg(Wt + (size_t)(k + 2) * N + j), w3 = __ldg(Wt + (size_t)(k + 3) * N + j);
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
          const double *hs = h + (size_t)s * stride + k;
          p[s][0] = fma(w0, hs[0], p[s][0]);
          p[s][1] = fma(w1, hs[1], p[s][1]);
          p[s][2] = fma(w2, hs[2], p[s][2]);
          p[s][3] = fma(w3, hs[3], p[s][3]);
        }
      }
      for (; k < Kin; ++k) {
        const double w = __ldg(Wt + (size_t)k * N + j);
#pragma unroll
        for (int s = 0; s < SPT; ++s) p[s][0] = fma(w, h[(size_t)s * stride + k], p[s][0]);
      }
      const double bj = __ldg(net.b[l] + j);
#pragma unroll
      for (int s = 0; s < SPT; ++s) {
        const double y = bj + ((p[s][0] + p[s][1]) + (p[s][2] + p[s][3]));
        hout[(size_t)(g * SPT + s) * stride + j] = last ? y : ampc_act<double>(net.act, y);
      }
    }
    __syncthreads();
    double *t2 = hin; hin = hout; hout = t2;
  }
  return hin;
}

__device__ __forceinline__ const double *ampc_mlp_f64_forward(const AmpcMlpF64 &net, double *h0, double *h1,
                                                              double *, int tid, int nthr) {
  return ampc_mlp_f64_forward_batch(net, 1, h0, h1, net.max_width, tid, nthr);
}

// Batched forward + input Jacobian of the raw network output w.r.t. the RAW
// (un-normalised) input, for m samples processed together by one CTA.
// Sample s uses  h0/h1/g at  +s*hstride  and Jacobian panels J0/J1 at +s*jstride
// (each panel max_width*nin).  On return *Jout (+s*jstride) is the panel
// [n_out][nin] = W_L D_{L-1} ... D_1 W_1 diag(1/xu_std) and the returned buffer
// (+s*hstride) holds the raw outputs.
__device__ __forceinline__ const double *ampc_mlp_f64_forward_jac_batch(const AmpcMlpF64 &net, int m, double *h0,
                                                                        double *h1, double *g, int hstride,
                                                                        double *J0, double *J1, int jstride,
                                                                        const double **Jout, int tid, int nthr) {
  const int nin = net.dims[0];
  double *hin = h0, *hout = h1, *Jp = J0, *Jn = J1;
  for (int l = 0; l < net.n_layers; ++l) {
    const int Kin = net.dims[l], N = net.dims[l + 1];
    const bool last = (l == net.n_layers - 1);
    for (int t = tid; t < m * N; t += nthr) {
      const int s = t / N, j = t - s * N;
      const double y = __ldg(net.b[l] + j) + ampc_dot_col(net.Wt[l], N, j, hin + (size_t)s * hstride, Kin);
      hout[(size_t)s * hstride + j] = last ? y : ampc_act<double>(net.act, y);
      g[(size_t)s * hstride + j] = last ? 1.0 : ampc_act_grad<double>(net.act, y);
    }
    __syncthreads();
    if (l == 0) {
      const int per = N * nin;
      for (int t = tid; t < m * per; t += nthr) {
        const int s = t / per, r = t - s * per;
        const int c = r / N, j = r - c * N;   // j fastest: coalesced weight reads
        const double v = __ldg(net.Wt[0] + (size_t)c * N + j) / __ldg(net.xu_std + c);
        Jn[(size_t)s * jstride + (size_t)j * nin + c] = v * g[(size_t)s * hstride + j];
      }
    } else {
      // J_l[j][c] = g[j] * sum_k W_l[k][j] J_{l-1}[k][c]: a thread owns output neuron j for FOUR input columns, so a
      // weight read feeds four multiply-adds and eight accumulators are in flight (even / odd k per column -- the
      // summation order per element is the one-column loop's: sum over even k, sum over odd k, tail into the even one,
      // even + odd).  A ragged last group recomputes the last column and drops the duplicates at the store.
      constexpr int CB = 4;
      const int ncg = (nin + CB - 1) / CB, per = N * ncg;
      const double *__restrict__ Wt = net.Wt[l];
      for (int t = tid; t < m * per; t += nthr) {
        const int s = t / per, r = t - s * per;
        const int cg = r / N, j = r - cg * N;   // j fastest: coalesced weight reads, broadcast panel reads
        const double *jp = Jp + (size_t)s * jstride;
        int cq[CB];
#pragma unroll
        for (int q = 0; q < CB; ++q) cq[q] = min(cg * CB + q, nin - 1);
        double p0[CB], p1[CB];
#pragma unroll
        for (int q = 0; q < CB; ++q) p0[q] = p1[q] = 0.0;
        int k = 0;
        for (; k + 2 <= Kin; k += 2) {
          const double w0 = __ldg(Wt + (size_t)k * N + j), w1 = __ldg(Wt + (size_t)(k + 1) * N + j);
          const double *r0 = jp + (size_t)k * nin, *r1 = r0 + nin;
#pragma unroll
          for (int q = 0; q < CB; ++q) {
            p0[q] = fma(w0, r0[cq[q]], p0[q]);
            p1[q] = fma(w1, r1[cq[q]], p1[q]);
          }
        }
        if (k < Kin) {
          const double w0 = __ldg(Wt + (size_t)k * N + j);
#pragma unroll
          for (int q = 0; q < CB; ++q) p0[q] = fma(w0, jp[(size_t)k * nin + cq[q]], p0[q]);
        }
        const double gj = g[(size_t)s * hstride + j];
        double *o = Jn + (size_t)s * jstride + (size_t)j * nin;
#pragma unroll
        for (int q = 0; q < CB; ++q)
          if (cg * CB + q < nin) o[cg * CB + q] = (p0[q] + p1[q]) * gj;
      }
    }
    __syncthreads();
    double *t2 = hin; hin = hout; hout = t2;
    double *t3 = Jp; Jp = Jn; Jn = t3;
  }
  *Jout = Jp;
  return hin;
}

__device__ __forceinline__ const double *ampc_mlp_f64_forward_jac(const AmpcMlpF64 &net, double *h0, double *h1,
                                                                  double *g, double *J0, double *J1,
                                                                  const double **Jout, int tid, int nthr) {
  return ampc_mlp_f64_forward_jac_batch(net, 1, h0, h1, g, net.max_width, J0, J1, net.max_width * net.dims[0], Jout,
                                        tid, nthr);
}
