// dist.cu -- row-partitioned CSR + CG across GPUs (placeholder, replaced by the real implementation).
#include "common.cuh"
extern "C" {
ViennaCLStatus ViennaCLCUDADdist_csr_create(ViennaCLBackend b, long long, long long, long long, ViennaCLInt, const unsigned int *, const unsigned int *, const double *, ViennaCLB200DistCsr *)
{ VCL_CHECK_BACKEND(b); return vcl_fail(b, ViennaCLGenericFailure, "dist_csr not built yet", __FILE__, __LINE__); }
ViennaCLStatus ViennaCLCUDADdist_csr_destroy(ViennaCLBackend b, ViennaCLB200DistCsr *) { VCL_CHECK_BACKEND(b); return ViennaCLSuccess; }
ViennaCLStatus ViennaCLCUDADdist_csrmv(ViennaCLBackend b, ViennaCLB200DistCsr, const double *, double *)
{ VCL_CHECK_BACKEND(b); return vcl_fail(b, ViennaCLGenericFailure, "dist_csr not built yet", __FILE__, __LINE__); }
ViennaCLStatus ViennaCLCUDADdist_csr_cg(ViennaCLBackend b, ViennaCLB200DistCsr, const double *, double *, ViennaCLB200SolverTag *)
{ VCL_CHECK_BACKEND(b); return vcl_fail(b, ViennaCLGenericFailure, "dist_csr not built yet", __FILE__, __LINE__); }
}
