// blas1.cuh -- internal BLAS-1 helpers shared with the solver drivers.
#pragma once
#include "common.cuh"

// <x,y> into out_dev[0], asynchronous on the backend stream.
ViennaCLStatus vcl_dot_async(ViennaCLBackend b, long long n, const double *x, int offx, int incx,
                             const double *y, int offy, int incy, double *out_dev);
// <x,y> to the host (synchronises the stream).
ViennaCLStatus vcl_dot_host(ViennaCLBackend b, long long n, const double *x, int offx, int incx,
                            const double *y, int offy, int incy, double *result);
