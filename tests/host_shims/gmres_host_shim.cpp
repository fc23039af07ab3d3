// Test shim (CPU): exposes the host half of a GMRES restart cycle (viennacl-dev_b200/csrc/gmres_host.cuh, plain C++) through a C entry
// point so that tests/test_host_logic.py can compare it with a numpy restatement of gmres.hpp:306-352.  Not part of the product library.
#include <vector>
#include "gmres_host.cuh"

struct Tag { double tolerance; int iters; };

extern "C" int gmres_cycle_host_shim(int k, const double *R, const double *xi, double tolerance, double rho_0, double norm_rhs, int per_iteration_stop,
                                     double *rho_inout, double *coef_out, int *iters_out, int *converged_out)
{
  Tag tag = {tolerance, 0};
  std::vector<double> hR(R, R + (size_t)k * k), vxi(xi, xi + k), eta(k), coef(k, 0.0);
  double rho = *rho_inout;
  bool conv = false;
  const size_t kk = vcl_gmres_cycle_host<double>(&tag, k, hR, vxi, eta, coef, rho, rho_0, norm_rhs, per_iteration_stop != 0, &conv);
  for (int i = 0; i < k; ++i) coef_out[i] = coef[i];
  *rho_inout = rho; *iters_out = tag.iters; *converged_out = conv ? 1 : 0;
  return (int)kk;
}
