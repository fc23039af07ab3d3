// gmres_launch.cuh -- small kernels and launch helpers of the GMRES drivers, shared by the single-domain driver (solvers.cu) and the
// row-partitioned one (dist.cu).
#pragma once
#include <algorithm>
#include "fused_kernels.cuh"

namespace VCL_NS
{
static __global__ void __launch_bounds__(VEC_THREADS)
scale_residual_kernel(long long n, real *res, real rho0)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    res[i] = res[i] / rho0;
}

static __global__ void __launch_bounds__(VEC_THREADS)
residual_kernel(long long n, real *res, const real *rhs)     // res = rhs - res
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    res[i] = rhs[i] - res[i];
}

static ViennaCLStatus launch_gs1(ViennaCLBackend b, int grid, const real *basis, long long n, long long isz, int k, real *out_h, int stride)
{
  VCL_REQUIRE(b, (isz & 1) == 0 && (reinterpret_cast<uintptr_t>(basis) & 15u) == 0u,
              "Krylov basis must be 16-byte aligned with an even internal size (the reference pads vectors to 128 entries, forwards.h:385)");
  (void)grid;
  const int gy = (k + GS1_COLS - 1) / GS1_COLS;                      // column groups
  const int gx = std::max(1, std::min(vcl_div_up(n / 2, VEC_THREADS), std::min(std::max(b->sm_count * 8 / gy, b->sm_count), VCL_MAX_BLOCKS)));
  gmres_gs1_kernel<<<dim3(gx, gy), VEC_THREADS, 0, b->stream>>>(basis, n, isz, k, out_h, stride, VCL_PARTIALS(b), b->tickets);
  VCL_LAUNCHED(b, "gmres_gs1_kernel");
  return ViennaCLSuccess;
}

static int scalar_grid(ViennaCLBackend b, long long n)
{
  return (int)std::max(1LL, std::min((n + VEC_THREADS - 1) / VEC_THREADS, (long long)std::min(b->sm_count * 8, VCL_MAX_BLOCKS)));
}
} // namespace VCL_NS
