// solver_state.cuh -- device-resident scalar state of a running Krylov solve.
// The reference keeps these scalars on the host and pays one blocking D2H per iteration (cg.hpp:164-172,
// bicgstab.hpp:180-189); here they live in device memory, are advanced by the last block of the reducing kernel,
// and the host only looks at them once per batch of iterations.
#pragma once

#define VCL_GMRES_MAX_KRYLOV 64

enum { VCL_RUNNING = 0, VCL_CONVERGED = 1, VCL_MAXIT = 2, VCL_BREAKDOWN = 3, VCL_MONITOR_STOP = 4 };

struct SolverState
{
  double sums[8];         // fully reduced inner products of the current iteration (after allreduce when distributed)
  double alpha, beta, omega;
  double norm_rhs;        // ||b||
  double norm_rhs_sq;     // ||b||^2
  double tol, abs_tol;
  double est;             // latest relative residual estimate (what a monitor would be shown)
  double residual_norm;
  double ip_rr0;          // preconditioned BiCGStab: <r, r0*>
  int iters;              // iterations completed (reference: tag.iters())
  int done;               // VCL_RUNNING / VCL_CONVERGED / ...
  int maxit;
  int need_restart;       // preconditioned BiCGStab: breakdown detected on device
  int last_restart;
  int restart_every;
  int pad0, pad1;
};
