// solver_state.cuh -- device-resident scalar state of a running Krylov solve.
// The reference keeps these scalars on the host and pays one blocking D2H per iteration (cg.hpp:164-172,
// bicgstab.hpp:180-189); here they live in device memory, are advanced by the last block of the reducing kernel,
// and the host only looks at them once per batch of iterations.
#pragma once
#include "prec.cuh"

#define VCL_GMRES_MAX_KRYLOV 64

enum { VCL_RUNNING = 0, VCL_CONVERGED = 1, VCL_MAXIT = 2, VCL_BREAKDOWN = 3, VCL_MONITOR_STOP = 4 };

namespace VCL_NS
{
struct SolverState
{
  real sums[8];         // fully reduced inner products of the current iteration (after allreduce when distributed)
  real alpha, beta, omega;
  real norm_rhs;        // ||b||
  real norm_rhs_sq;     // ||b||^2
  real tol, abs_tol;
  real est;             // latest relative residual estimate (what a monitor would be shown)
  real residual_norm;
  real ip_rr0;          // preconditioned BiCGStab: <r, r0*>
  int iters;              // iterations completed (reference: tag.iters())
  int done;               // VCL_RUNNING / VCL_CONVERGED / ...
  int maxit;
  int need_restart;       // preconditioned BiCGStab: breakdown detected on device
  int last_restart;
  int restart_every;
  int pad0, pad1;
};
} // namespace VCL_NS
