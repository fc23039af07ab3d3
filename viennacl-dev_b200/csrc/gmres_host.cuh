// gmres_host.cuh -- the host half of one pipelined GMRES restart cycle (gmres.hpp:306-352): Krylov-space truncation on loss of
// orthogonality, the residual-norm recurrence with the per-iteration stopping rule of the preconditioned path (gmres.hpp:579-584),
// and the triangular solve for the update coefficients.  Shared by the single-domain driver (solvers.cu, both precisions) and the
// row-partitioned driver (dist.cu): R and xi are global quantities there, so every rank takes identical decisions.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

// k: inner iterations executed (leading dimension of hR is k); returns the number of basis vectors the update uses (kk).
// tag->iters is advanced as the reference does; rho is updated in place; *converged only ever set for per_iteration_stop.
template<class T, class Tag>
static inline size_t vcl_gmres_cycle_host(Tag *tag, int k, const std::vector<T> &hR, const std::vector<T> &xi, std::vector<T> &eta, std::vector<T> &coef,
                                          T &rho, T rho_0, T norm_rhs, bool per_iteration_stop, bool *converged)
{
  size_t kk = (size_t)k;
  const size_t full = kk;                                                      // gmres.hpp:306-314
  for (size_t i = 0; i < kk; ++i)
    if (std::fabs(hR[i + i * kk]) < tag->tolerance * hR[0]) { kk = i; break; }

  *converged = false;
  for (size_t i = 0; i < kk; ++i)                                              // gmres.hpp:318-331
  {
    tag->iters += 1;
    if (xi[i] >= rho || xi[i] <= -rho) { kk = i; break; }
    rho *= std::sin(std::acos(xi[i] / rho));
    if (per_iteration_stop && std::fabs(rho * rho_0 / norm_rhs) < tag->tolerance) { kk = i + 1; *converged = true; break; }   // gmres.hpp:579-584
  }

  eta = xi;                                                                    // gmres.hpp:336-345
  for (long i2 = (long)kk - 1; i2 > -1; --i2)
  {
    const size_t i = (size_t)i2;
    for (size_t j = i + 1; j < kk; ++j) eta[i] -= hR[i + j * full] * eta[j];
    eta[i] /= hR[i + i * full];
  }
  for (size_t i = 0; i < kk; ++i) coef[i] = rho_0 * eta[i];                    // gmres.hpp:351-352
  return kk;
}
