// Linear-model prediction on the device, float64.
//
// Replaces ARX.pred / pred_batch (autompc/sysid/arx.py:146-154) and Koopman.pred / pred_batch
// (autompc/sysid/koopman.py:165-173): both are  statesnew = (A @ states.T + B @ ctrls.T).T  with a state that is a
// history stack (ARX) or a lifted observation (Koopman).  SURVEY.md 8(f) row 4.
//
// Mapping: a CTA owns SB samples; thread i owns output rows i, i+NT, ... and keeps SB accumulators, so every element
// of A (stored transposed: neighbouring threads read neighbouring rows) is loaded once per SB samples; the samples'
// states and controls sit in shared memory and are read as broadcasts.  The op is memory-bound: per sample it moves
// (2 ns + nu) * 8 B of HBM traffic against 2 ns (ns + nu) flops; A and B stay in L2.
#include <vector>

#include "ampc_common.cuh"

struct ampc_linear {
  int ns = 0, nu = 0, device = 0;
  double *d_At = nullptr;   // [ns + nu][ns]: rows 0..ns-1 = A^T, rows ns.. = B^T
  double *d_scratch = nullptr;   // grow-only staging of X | U | Xn (no cudaMalloc per call)
  size_t scratch_doubles = 0;
};

namespace {
constexpr int NT = 128;
constexpr int SB = 8;

__global__ void __launch_bounds__(NT) linear_pred_batch_kernel(int ns, int nu, int batch, const double *__restrict__ At,
                                                               const double *__restrict__ X, const double *__restrict__ U,
                                                               double *__restrict__ Xn) {
  extern __shared__ double sm_lin[];          // [SB][ns + nu]
  const int s0 = blockIdx.x * SB, nin = ns + nu;
  for (int t = threadIdx.x; t < SB * nin; t += NT) {
    const int q = t / nin, j = t - q * nin, s = s0 + q;
    sm_lin[t] = s < batch ? (j < ns ? X[(size_t)s * ns + j] : U[(size_t)s * nu + (j - ns)]) : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ns; i += NT) {
    double acc[SB];
#pragma unroll
    for (int q = 0; q < SB; ++q) acc[q] = 0.0;
    for (int j = 0; j < nin; ++j) {
      const double a = __ldg(At + (size_t)j * ns + i);
#pragma unroll
      for (int q = 0; q < SB; ++q) acc[q] = fma(a, sm_lin[q * nin + j], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < SB; ++q)
      if (s0 + q < batch) Xn[(size_t)(s0 + q) * ns + i] = acc[q];
  }
}
}  // namespace

extern "C" int ampc_linear_create(ampc_linear **out, int32_t ns, int32_t nu, const double *A, const double *B,
                                  int32_t device) {
  AMPC_REQUIRE(out && A && B, AMPC_ERR_INVALID, "null argument");
  *out = nullptr;
  AMPC_REQUIRE(ns >= 1 && nu >= 1 && (size_t)SB * (ns + nu) * sizeof(double) <= 200 * 1024, AMPC_ERR_INVALID,
               "bad linear-model dims ns=%d nu=%d", ns, nu);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  AMPC_REQUIRE(ce == cudaSuccess && ndev > 0, AMPC_ERR_CUDA, "no CUDA device: libampc_b200 has no CPU fallback (%s)",
               cudaGetErrorString(ce));
  AMPC_REQUIRE(device >= 0 && device < ndev, AMPC_ERR_INVALID, "device %d of %d", device, ndev);
  AMPC_CUDA_CHECK(cudaSetDevice(device));
  std::vector<double> at((size_t)(ns + nu) * ns);
  for (int i = 0; i < ns; ++i) {
    for (int j = 0; j < ns; ++j) at[(size_t)j * ns + i] = A[(size_t)i * ns + j];
    for (int j = 0; j < nu; ++j) at[(size_t)(ns + j) * ns + i] = B[(size_t)i * nu + j];
  }
  ampc_linear *h = new ampc_linear();
  h->ns = ns; h->nu = nu; h->device = device;
  cudaError_t e = cudaMalloc(&h->d_At, at.size() * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_At, at.data(), at.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = ampc_raise_smem_limit((const void *)linear_pred_batch_kernel, (size_t)SB * (ns + nu) * sizeof(double));
  if (e != cudaSuccess) {
    ampc_set_error("linear model create: %s", cudaGetErrorString(e));
    cudaFree(h->d_At);
    delete h;
    return AMPC_ERR_CUDA;
  }
  *out = h;
  return AMPC_OK;
}

extern "C" int ampc_linear_destroy(ampc_linear *h) {
  if (!h) return AMPC_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_At);
  cudaFree(h->d_scratch);
  delete h;
  return AMPC_OK;
}

extern "C" int ampc_linear_pred_batch(ampc_linear *h, int32_t batch, const double *X, const double *U, double *Xn) {
  AMPC_REQUIRE(h && X && U && Xn && batch >= 0, AMPC_ERR_INVALID, "bad argument");
  if (batch == 0) return AMPC_OK;
  AMPC_CUDA_CHECK(cudaSetDevice(h->device));
  const size_t nX = (size_t)batch * h->ns, nU = (size_t)batch * h->nu;
  if (2 * nX + nU > h->scratch_doubles) {
    cudaFree(h->d_scratch);
    h->d_scratch = nullptr;
    h->scratch_doubles = 0;
    AMPC_CUDA_CHECK(cudaMalloc(&h->d_scratch, (2 * nX + nU) * sizeof(double)));
    h->scratch_doubles = 2 * nX + nU;
  }
  double *d = h->d_scratch;
  double *dX = d, *dU = dX + nX, *dXn = dU + nU;
  cudaError_t e = cudaMemcpy(dX, X, nX * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dU, U, nU * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    linear_pred_batch_kernel<<<(batch + SB - 1) / SB, NT, (size_t)SB * (h->ns + h->nu) * sizeof(double)>>>(
        h->ns, h->nu, batch, h->d_At, dX, dU, dXn);
    ampc_count_launch();
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(Xn, dXn, nX * sizeof(double), cudaMemcpyDeviceToHost);
  AMPC_CUDA_CHECK(e);
  return AMPC_OK;
}
