// blas1.cu -- the BLAS-1 subset the solver drivers reach outside their fused loops.
// Replaces cuda/vector_operations.hpp:77 (av), :179 (avbv), :483 (avbv_v), :782 (vector_assign), :870 (element_op '/'),
// :1273-1579 (inner_prod, inner_prod_cpu), :2018-2448 (norm_2, norm_2_cpu).  Unlike the reference, reductions do not
// cudaMalloc/cudaFree a temporary per call (cuda/vector_operations.hpp:1557-1578): scratch lives in the backend handle.
#include "common.cuh"
#include "blas1.cuh"
#include <algorithm>

namespace VCL_NS
{

template<int OP>   // 0: x = a*y ; 1: x = a*y + b*z ; 2: x += a*y + b*z ; 3: x = value ; 4: x = y / z
__global__ void __launch_bounds__(256)
ew_kernel(long long n, real *x, int offx, int incx, const real *y, int offy, int incy, real a,
          const real *z, int offz, int incz, real bb)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
  {
    const size_t ix = (size_t)i * incx + offx;
    if (OP == 0) x[ix] = a * y[(size_t)i * incy + offy];
    if (OP == 1) x[ix] = a * y[(size_t)i * incy + offy] + bb * z[(size_t)i * incz + offz];
    if (OP == 2) x[ix] += a * y[(size_t)i * incy + offy] + bb * z[(size_t)i * incz + offz];
    if (OP == 3) x[ix] = a;
    if (OP == 4) x[ix] = y[(size_t)i * incy + offy] / z[(size_t)i * incz + offz];
  }
}

static int ew_grid(ViennaCLBackend b, long long n) { return (int)std::max(1LL, std::min((n + 255) / 256, (long long)b->sm_count * 8)); }

template<int OP>
static ViennaCLStatus ew_launch(ViennaCLBackend b, long long n, real *x, int offx, int incx, const real *y, int offy, int incy, real a,
                                const real *z, int offz, int incz, real bb)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, n >= 0, "negative size");
  if (n == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, x != nullptr, "null pointer");
  ew_kernel<OP><<<ew_grid(b, n), 256, 0, b->stream>>>(n, x, offx, incx, y, offy, incy, a, z, offz, incz, bb);
  VCL_LAUNCHED(b, "ew_kernel");
  return ViennaCLSuccess;
}

// <x,y> (y == x gives the squared 2-norm): deterministic two-stage sum, result left in out[0] on the device.
__global__ void __launch_bounds__(256)
dot_kernel(long long n, const real *x, int offx, int incx, const real *y, int offy, int incy,
           real *partials, unsigned int *ticket, real *out)
{
  __shared__ real s_red[32];
  real acc[1] = {0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc[0] = fma(x[(size_t)i * incx + offx], y[(size_t)i * incy + offy], acc[0]);
  if (grid_sum_last_block<1>(acc, partials, ticket, s_red) && threadIdx.x == 0) out[0] = acc[0];
}

ViennaCLStatus vcl_dot_async(ViennaCLBackend b, long long n, const real *x, int offx, int incx,
                             const real *y, int offy, int incy, real *out_dev)
{
  int grid = (int)std::max(1LL, std::min((n + 255) / 256, (long long)std::min(b->sm_count * 8, VCL_MAX_BLOCKS)));
  dot_kernel<<<grid, 256, 0, b->stream>>>(n, x, offx, incx, y, offy, incy, VCL_PARTIALS(b), b->tickets, out_dev);
  VCL_LAUNCHED(b, "dot_kernel");
  return ViennaCLSuccess;
}

ViennaCLStatus vcl_dot_host(ViennaCLBackend b, long long n, const real *x, int offx, int incx,
                            const real *y, int offy, int incy, real *result)
{
  if (n == 0) { *result = 0.0; return ViennaCLSuccess; }
  VCL_TRY(vcl_dot_async(b, n, x, offx, incx, y, offy, incy, VCL_DSCAL(b)));
  VCL_CUDA(b, cudaMemcpyAsync(VCL_HSCAL(b), VCL_DSCAL(b), sizeof(real), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  *result = VCL_HSCAL(b)[0];
  return ViennaCLSuccess;
}

extern "C" {

ViennaCLStatus ViennaCLCUDADav(ViennaCLBackend b, ViennaCLInt n, real *x, ViennaCLInt offx, ViennaCLInt incx,
                               const real *y, ViennaCLInt offy, ViennaCLInt incy, real alpha)
{ return ew_launch<0>(b, n, x, offx, incx, y, offy, incy, alpha, nullptr, 0, 1, 0.0); }

ViennaCLStatus ViennaCLCUDADavbv(ViennaCLBackend b, ViennaCLInt n, real *x, ViennaCLInt offx, ViennaCLInt incx,
                                 const real *y, ViennaCLInt offy, ViennaCLInt incy, real alpha,
                                 const real *z, ViennaCLInt offz, ViennaCLInt incz, real beta)
{ return ew_launch<1>(b, n, x, offx, incx, y, offy, incy, alpha, z, offz, incz, beta); }

ViennaCLStatus ViennaCLCUDADavbv_v(ViennaCLBackend b, ViennaCLInt n, real *x, ViennaCLInt offx, ViennaCLInt incx,
                                   const real *y, ViennaCLInt offy, ViennaCLInt incy, real alpha,
                                   const real *z, ViennaCLInt offz, ViennaCLInt incz, real beta)
{ return ew_launch<2>(b, n, x, offx, incx, y, offy, incy, alpha, z, offz, incz, beta); }

ViennaCLStatus ViennaCLCUDADassign(ViennaCLBackend b, ViennaCLInt n, real *x, ViennaCLInt offx, ViennaCLInt incx, real value)
{ return ew_launch<3>(b, n, x, offx, incx, nullptr, 0, 1, value, nullptr, 0, 1, 0.0); }

ViennaCLStatus ViennaCLCUDADelement_div(ViennaCLBackend b, ViennaCLInt n, real *x, ViennaCLInt offx, ViennaCLInt incx,
                                        const real *y, ViennaCLInt offy, ViennaCLInt incy,
                                        const real *z, ViennaCLInt offz, ViennaCLInt incz)
{ return ew_launch<4>(b, n, x, offx, incx, y, offy, incy, 0.0, z, offz, incz, 0.0); }

ViennaCLStatus ViennaCLCUDADdot(ViennaCLBackend b, ViennaCLInt n, real *result_host,
                                const real *x, ViennaCLInt offx, ViennaCLInt incx,
                                const real *y, ViennaCLInt offy, ViennaCLInt incy)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, n >= 0 && result_host, "bad arguments");
  return vcl_dot_host(b, n, x, offx, incx, y, offy, incy, result_host);
}

ViennaCLStatus ViennaCLCUDADnrm2(ViennaCLBackend b, ViennaCLInt n, real *result_host,
                                 const real *x, ViennaCLInt offx, ViennaCLInt incx)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, n >= 0 && result_host, "bad arguments");
  real s = 0.0;
  VCL_TRY(vcl_dot_host(b, n, x, offx, incx, x, offx, incx, &s));
  *result_host = sqrt(s);     // final sqrt on the host, as norm_2_cpu does (cuda/vector_operations.hpp:2431-2448)
  return ViennaCLSuccess;
}

} // extern "C"
} // namespace VCL_NS
