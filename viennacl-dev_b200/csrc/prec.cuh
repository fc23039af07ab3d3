// prec.cuh -- precision of this translation unit.  The kernel / driver sources (blas1, spmv, solvers, gen) are compiled
// twice: as they are (`real` = double, entry points ViennaCLCUDAD...) and with -DVCL_F32 (`real` = float, entry points
// ViennaCLCUDAS...; the names are mapped by prec_names_f32.h, generated together with include/vcl_b200_float.h by
// tools/gen_float_header.py).  Everything precision-dependent lives in namespace VCL_NS, so the two builds never share a
// definition.  The row-partitioned path (dist.cu, peer.cuh) exists in double only.
#pragma once
#include <cuda_runtime.h>

#ifdef VCL_F32
typedef float real;
typedef float2 real2;
#define VCL_NS vcl_f32
#define VCL_PTX_REAL "f32"
#define VCL_PTX_REG "f"
#include "prec_names_f32.h"
#else
typedef double real;
typedef double2 real2;
#define VCL_NS vcl_f64
#define VCL_PTX_REAL "f64"
#define VCL_PTX_REG "d"
#endif

#ifdef __CUDACC__
// rounded (never contracted) multiply / add: the two roundings of the reference's unfused CSR chain
__device__ __forceinline__ double rmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double radd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float rmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float radd(float a, float b) { return __fadd_rn(a, b); }
// v != 0 without the floating-point pipe (true for NaN, false for +-0)
__device__ __forceinline__ bool nonzero(double v) { return (__double_as_longlong(v) << 1) != 0; }
__device__ __forceinline__ bool nonzero(float v) { return (__float_as_uint(v) << 1) != 0u; }
#endif

// scratch of the backend handle, viewed in this build's precision
#define VCL_PARTIALS(b) (reinterpret_cast<real*>((b)->partials))
#define VCL_DSCAL(b)    (reinterpret_cast<real*>((b)->dscal))
#define VCL_HSCAL(b)    (reinterpret_cast<real*>((b)->hscal))
#define VCL_DSTATE(b)   (reinterpret_cast<SolverState*>((b)->dstate))
#define VCL_HSTATE(b)   (reinterpret_cast<SolverState*>((b)->hstate))
