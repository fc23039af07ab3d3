// persistent.cuh -- whole solver iterations inside ONE cooperative kernel (small and medium systems).
//
// For systems up to a few million rows a CG iteration is not bound by HBM but by fixed latencies: two dependent launches,
// two kernel ramps and two last-CTA reduction tails cost ~16 us per iteration whatever the size (measured on B200: 256^2
// 16.0 us, 1024^2 35.4 us per iteration).  Here one cooperative launch (one CTA set resident on every SM) runs up to
// `iterations` iterations: the same two bodies as the stand-alone kernels -- cg_update_body, then the CSR row-block
// product with the fused inner products whose last CTA advances alpha / beta / the convergence test in device memory --
// separated by grid-wide barriers instead of kernel boundaries.  Arithmetic, reduction order within a CTA and the
// stopping rule are those of the two-kernel path; only the number of CTAs that share the vector update differs.
#pragma once
#include <cooperative_groups.h>
#include "fused_kernels.cuh"
#include "spmv_kernels.cuh"

namespace VCL_NS
{
namespace cgrp = cooperative_groups;

__global__ void __launch_bounds__(CSR_BLOCK_THREADS, CSR_MIN_CTAS)
cg_persistent_kernel(CsrDev A, XVec xv, long long n, real *x, real *p, real *r, real *Ap,
                     SolverState *st, real *partials, unsigned int *ticket, int iterations)
{
  cgrp::grid_group grid = cgrp::this_grid();
  const PushRanges no_push = PushRanges();
  for (int it = 0; it < iterations; ++it)
  {
    // st->done is written by ONE thread of the grid before the barrier that ends an iteration: uniform for all CTAs
    if (*reinterpret_cast<volatile int*>(&st->done) != VCL_RUNNING) break;
    cg_update_body(n, x, p, r, Ap, 0.0, 0.0, st, partials, ticket, &st->sums[0], no_push);
    grid.sync();                                           // new p (and <r,r>) visible to every CTA
    EpiFused<STEP_CG, false, false> epi = {Ap, p, nullptr, nullptr, partials, ticket, st, &st->sums[1], &st->sums[2], nullptr, {0.0, 0.0, 0.0}, nullptr};
    csr_stream_body<EpiFused<STEP_CG, false, false>, false, true>(A, xv, epi);
    grid.sync();                                           // Ap, alpha, beta, done visible
  }
}
}
