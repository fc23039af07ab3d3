// mixed.cu -- mixed-precision CG (linalg/mixed_precision_cg.hpp:95-186): double outside, float inside.
//
// The reference keeps a float copy of the matrix and runs a classical CG on it (SpMV + 2 blocking inner products + 3
// vector kernels per iteration); whenever the float residual has dropped by `inner_tol` it folds the float iterate into the
// double result, recomputes the residual in double and restarts the float iteration from it.  A restarted CG on
// A e = r with e_0 = 0 is exactly a fresh solve, so here every inner phase is ONE call of the fused single-precision
// pipelined CG (ViennaCLCUDAScsr_cg: 2 kernels per iteration, 8 instead of 12 bytes per matrix entry, scalars on the
// device) with tolerance sqrt(inner_tol) -- the reference compares squared norms (:160) -- and the remaining iteration
// budget; the outer step (x += e, r = b - A x, ||r||) runs in double.  Compiled once (no -DVCL_F32 twin).
#include "common.cuh"
#include "vcl_b200_float.h"
#include <algorithm>
#include <cmath>

namespace
{
__global__ void __launch_bounds__(256) d2s_kernel(long long n, const double * __restrict__ x, float * __restrict__ y)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = (float)x[i];
}
__global__ void __launch_bounds__(256) s2d_kernel(long long n, const float * __restrict__ x, double * __restrict__ y)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = (double)x[i];
}
// result += e (float -> double, mixed_precision_cg.hpp:162-163) and r = b, the start value of r = b - A result
__global__ void __launch_bounds__(256) fold_kernel(long long n, double * __restrict__ result, const float * __restrict__ e,
                                                   const double * __restrict__ b, double * __restrict__ r)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
  {
    result[i] += (double)e[i];
    r[i] = b[i];
  }
}
int grid_for(ViennaCLBackend b, long long n) { return (int)std::max(1LL, std::min((n + 255) / 256, (long long)b->sm_count * 8)); }
size_t pad256(size_t bytes) { return (bytes + 255) / 256 * 256; }
}

extern "C" ViennaCLStatus ViennaCLCUDAconvert_DtoS(ViennaCLBackend b, long long n, const double *x, float *y)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, n >= 0 && (n == 0 || (x && y)), "bad arguments");
  if (n == 0) return ViennaCLSuccess;
  d2s_kernel<<<grid_for(b, n), 256, 0, b->stream>>>(n, x, y);
  VCL_LAUNCHED(b, "d2s_kernel");
  return ViennaCLSuccess;
}

extern "C" ViennaCLStatus ViennaCLCUDAconvert_StoD(ViennaCLBackend b, long long n, const float *x, double *y)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, n >= 0 && (n == 0 || (x && y)), "bad arguments");
  if (n == 0) return ViennaCLSuccess;
  s2d_kernel<<<grid_for(b, n), 256, 0, b->stream>>>(n, x, y);
  VCL_LAUNCHED(b, "s2d_kernel");
  return ViennaCLSuccess;
}

extern "C" ViennaCLStatus ViennaCLCUDADcsr_mixed_precision_cg(ViennaCLBackend b, const ViennaCLCUDADcsr *A, const float *values_float,
                                                              const double *rhs, double *x, float inner_tolerance, ViennaCLB200SolverTag *tag)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, A && tag, "null matrix / tag");
  VCL_REQUIRE(b, A->rows == A->cols && A->rows >= 0 && A->nnz >= 0, "mixed-precision CG needs a square matrix");
  VCL_REQUIRE(b, inner_tolerance > 0.0f && inner_tolerance < 1.0f, "inner tolerance must lie in (0, 1)");
  const long long n = A->rows;
  tag->iters = 0; tag->error = 0.0;
  if (n == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, rhs && x, "null vector");
  VCL_CUDA(b, cudaSetDevice(b->device));

  // workspace: [ what the float CG carves for itself (3 float vectors) | r (double) | r_low | e_low | values_low ]
  const size_t inner = 3 * pad256(sizeof(float) * (size_t)n);
  const size_t need = inner + pad256(sizeof(double) * (size_t)n) + 2 * pad256(sizeof(float) * (size_t)n)
                    + (values_float ? 0 : pad256(sizeof(float) * (size_t)std::max(A->nnz, 1)));
  VCL_TRY(vcl_ws_reserve(b, need));
  char *base = static_cast<char*>(b->ws) + inner;
  double *r = reinterpret_cast<double*>(base);             base += pad256(sizeof(double) * (size_t)n);
  float *r_low = reinterpret_cast<float*>(base);           base += pad256(sizeof(float) * (size_t)n);
  float *e_low = reinterpret_cast<float*>(base);           base += pad256(sizeof(float) * (size_t)n);
  const float *va_low = values_float;
  if (!va_low)
  {
    float *conv = reinterpret_cast<float*>(base);
    VCL_TRY(ViennaCLCUDAconvert_DtoS(b, A->nnz, A->values, conv));                   // mixed_precision_cg.hpp:131-134
    va_low = conv;
  }
  ViennaCLCUDAScsr A_low = {A->rows, A->cols, A->nnz, A->row_ptr, A->col_idx, va_low, A->row_blocks, A->num_blocks};   // index arrays shared

  double nrm = 0.0;
  VCL_TRY(ViennaCLCUDADnrm2(b, (ViennaCLInt)n, &nrm, rhs, 0, 1));
  const double norm_rhs_squared = nrm * nrm;                                          // :110-112
  VCL_CUDA(b, cudaMemsetAsync(x, 0, sizeof(double) * n, b->stream));
  if (norm_rhs_squared <= 0) return ViennaCLSuccess;                                  // :114-115
  VCL_CUDA(b, cudaMemcpyAsync(r, rhs, sizeof(double) * n, cudaMemcpyDeviceToDevice, b->stream));

  double new_ip_rr = 0.0;
  int total = 0;
  while (total < tag->max_iterations)
  {
    VCL_TRY(ViennaCLCUDAconvert_DtoS(b, n, r, r_low));                                // :127-128 / :172-175
    ViennaCLB200SolverTagS inner_tag;
    inner_tag.tolerance = std::sqrt((double)inner_tolerance); inner_tag.abs_tolerance = 0.0;
    inner_tag.max_iterations = tag->max_iterations - total; inner_tag.krylov_dim = 0; inner_tag.max_iterations_before_restart = 0;
    inner_tag.precond = ViennaCLB200PrecondNone; inner_tag.monitor = nullptr; inner_tag.monitor_user = nullptr;
    inner_tag.iters = 0; inner_tag.error = 0.0;
    VCL_TRY(ViennaCLCUDAScsr_cg(b, &A_low, r_low, e_low, &inner_tag));
    total += std::max(inner_tag.iters, 1);
    // result += e; r = b - A result, in double (:162-166)
    fold_kernel<<<grid_for(b, n), 256, 0, b->stream>>>(n, x, e_low, rhs, r);
    VCL_LAUNCHED(b, "fold_kernel");
    VCL_TRY(ViennaCLCUDADcsrmv(b, A->rows, A->cols, A->nnz, A->row_ptr, A->col_idx, A->values, A->row_blocks, A->num_blocks,
                               x, 0, 1, -1.0, r, 0, 1, 1.0));
    VCL_TRY(ViennaCLCUDADnrm2(b, (ViennaCLInt)n, &nrm, r, 0, 1));
    new_ip_rr = nrm * nrm;
    if (tag->monitor && tag->monitor(x, std::sqrt(new_ip_rr / norm_rhs_squared), tag->monitor_user)) break;
    if (new_ip_rr / norm_rhs_squared < tag->tolerance * tag->tolerance) break;        // :168-169
    if (inner_tag.iters == 0) break;                                                  // the float iteration cannot make progress any more
  }
  tag->iters = total;                                                                  // :151
  tag->error = std::sqrt(new_ip_rr / norm_rhs_squared);                                // :182
  return ViennaCLSuccess;
}
