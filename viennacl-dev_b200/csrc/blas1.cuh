// blas1.cuh -- internal BLAS-1 helpers shared with the solver drivers.
#pragma once
#include "common.cuh"
#include "prec.cuh"

namespace VCL_NS
{

// <x,y> into out_dev[0], asynchronous on the backend stream.
ViennaCLStatus vcl_dot_async(ViennaCLBackend b, long long n, const real *x, int offx, int incx,
                             const real *y, int offy, int incy, real *out_dev);
// <x,y> to the host (synchronises the stream).
ViennaCLStatus vcl_dot_host(ViennaCLBackend b, long long n, const real *x, int offx, int incx,
                            const real *y, int offy, int incy, real *result);
} // namespace VCL_NS
