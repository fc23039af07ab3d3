// persistent.cuh -- whole solver iterations inside ONE cooperative kernel (small and medium systems).
//
// For systems up to a few million rows a CG iteration is not bound by HBM but by fixed latencies: two dependent launches,
// two kernel ramps and two last-CTA reduction tails cost ~16 us per iteration whatever the size (measured on B200: 256^2
// 16.0 us, 1024^2 35.4 us per iteration).  Here one cooperative launch (one CTA set resident on every SM) runs up to
// `iterations` iterations.  Per iteration:
//     vector update (x += alpha p; r -= alpha Ap; p = r + beta p)        -> one partial <r,r> per CTA
//     grid barrier
//     CSR row-block product Ap = A p (csr_stream_body, the stand-alone kernel's body) -> partial <Ap,Ap>, <p,Ap> per CTA
//     grid barrier
// and after each barrier EVERY CTA sums the per-CTA partials itself, in the same fixed order, and advances its own copy of
// the solver scalars (shared memory) -- no ticket, no "last CTA" tail, no second pass through global memory for alpha and
// beta; all CTAs take identical decisions because they perform identical arithmetic on identical data.  CTA 0 writes the
// state back for the host when the launch ends.  Arithmetic per entry and the stopping rule are those of the two-kernel
// path (cg.hpp:128-187); only the grouping of the partial sums differs.
// The file holds five kernels of this form: cg_persistent_kernel (below, two phases), cg_onepass_kernel (ONE phase and one barrier per
// iteration: the product recomputes the updated search direction on the fly; used up to 600 k rows), pcg_persistent_kernel (Jacobi /
// row scaling), bicgstab_persistent_kernel (pipelined BiCGStab, four phases) and gmres_persistent_kernel (one restart cycle); the drivers in
// solvers.cu choose them by system size (persistent_cg_wanted) and fall back to the multi-kernel form when a cooperative
// launch does not fit.
#pragma once
#include <cooperative_groups.h>
#include "fused_kernels.cuh"
#include "spmv_kernels.cuh"

namespace VCL_NS
{
namespace cgrp = cooperative_groups;

// epilogue of the product inside the persistent kernel: same per-row work as EpiFused<STEP_CG>, per-CTA partials only
struct EpiCgPartial
{
  real *Ap; const real *p; real *partials;
  real acc[2];
  static constexpr int NQ = 2;
  static constexpr bool COO = false;
  typedef real Pre;
  __device__ __forceinline__ real init(real) const { return 0.0; }
  __device__ __forceinline__ real term_scale() const { return 1.0; }
  __device__ __forceinline__ bool skip() const { return false; }
  __device__ __forceinline__ real pre(u32 r) const { return p[r]; }
  __device__ __forceinline__ void row(u32 r, real dot, real p_r)
  {
    Ap[r] = dot;
    acc[0] = fma(dot, dot, acc[0]);
    acc[1] = fma(p_r, dot, acc[1]);
  }
  __device__ __forceinline__ void finish(real *smem)
  {
    block_sum<2>(acc, smem);
    if (threadIdx.x == 0) { partials[blockIdx.x] = acc[0]; partials[VCL_MAX_BLOCKS + blockIdx.x] = acc[1]; }
  }
};

// totals of NQ partial arrays (stride VCL_MAX_BLOCKS), summed by this CTA in a fixed order; valid in thread 0
template<int NQ>
__device__ __forceinline__ void sum_partials(const real *partials, real (&tot)[NQ], real *smem)
{
#pragma unroll
  for (int q = 0; q < NQ; ++q)
  {
    tot[q] = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) tot[q] += __ldcg(partials + q * VCL_MAX_BLOCKS + i);
  }
  block_sum<NQ>(tot, smem);
}

// the same for NQ arrays given by pointer (arrays that are double-buffered separately), ONE block-level reduction for all of them
template<int NQ>
__device__ __forceinline__ void sum_partials_of(const real * const (&arr)[NQ], real (&tot)[NQ], real *smem)
{
#pragma unroll
  for (int q = 0; q < NQ; ++q)
  {
    tot[q] = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) tot[q] += __ldcg(arr[q] + i);
  }
  block_sum<NQ>(tot, smem);
}

__global__ void __launch_bounds__(CSR_BLOCK_THREADS, CSR_MIN_CTAS)
cg_persistent_kernel(CsrDev A, XVec xv, long long n, real *x, real *p, real *r, real *Ap,
                     SolverState *st, real *partials, int iterations)
{
  cgrp::grid_group grid = cgrp::this_grid();
  __shared__ SolverState s_st;                             // this CTA's copy of the scalars
  __shared__ real s_sum[3 * 32];
  if (threadIdx.x == 0) s_st = *st;
  __syncthreads();
  CsrCarry carry = {0u, 0};                                // descriptors / barriers / first copy of the product carried between iterations

  for (int it = 0; it < iterations; ++it)
  {
    if (s_st.done != VCL_RUNNING) break;                   // identical in every CTA
    // partial arrays: [0]: <Ap,Ap>, [1]: <p,Ap>, [2] / [3]: <r,r> of even / odd iterations.  <r,r> is double-buffered because
    // it is written before the first barrier of an iteration but read after the second one: a CTA that is already in
    // the next iteration's update must not overwrite what a slower CTA is still summing.
    real *part_rr = partials + (2 + (it & 1)) * VCL_MAX_BLOCKS;
    // ---- vector update with this CTA's share of the entries (the loop of cg_update_kernel) ----
    {
      real acc[1] = {cg_update_entries<false>(n, x, p, r, Ap, s_st.alpha, s_st.beta, PushRanges())};
      block_sum<1>(acc, s_sum);
      if (threadIdx.x == 0) part_rr[blockIdx.x] = acc[0];
    }
    grid.sync();                                           // new p and all <r,r> partials visible
    // ---- Ap = A p with the fused inner products ----
    EpiCgPartial epi = {Ap, p, partials, {0.0, 0.0}};
    csr_stream_body<EpiCgPartial, false, true>(A, xv, epi, &carry);
    grid.sync();                                           // Ap and all partials visible
    real tot[3];
    const real * const arrs[3] = {part_rr, partials, partials + VCL_MAX_BLOCKS};
    sum_partials_of<3>(arrs, tot, s_sum);                  // <r,r>, <Ap,Ap>, <p,Ap>: one block-level reduction for the three
    if (threadIdx.x == 0)
    {
      s_st.sums[0] = tot[0]; s_st.sums[1] = tot[1]; s_st.sums[2] = tot[2];
      cg_advance(&s_st);                                   // cg.hpp:170-180
    }
    __syncthreads();
  }
  {
    EpiCgPartial epi = {Ap, p, partials, {0.0, 0.0}};
    csr_stream_body<EpiCgPartial, false, true>(A, xv, epi, &carry, true);        // wait for the copy issued ahead, if any
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *st = s_st;
}
// ------------------------------------------------------------------------------------------------
// One-pass persistent CG: ONE phase and ONE grid barrier per iteration.
//
// The two-phase kernel above is bound by fixed latencies on BASELINE config 1 (1024^2: 25.9 us per iteration of which ~11 us do
// not depend on the size): update phase, barrier, product phase, barrier.  The barrier between update and product exists
// only because the product gathers p_new at OTHER rows.  Here the product recomputes those entries on the fly,
//     p_new[c] = fma(beta, p[c], fma(-alpha, Ap[c], r[c]))      -- bit-identical to what the update phase would have stored,
// from the PREVIOUS iteration's r / p / Ap (three gathers per entry instead of one; they hit L1 / L2), and the row's own update
// (x += alpha p; r_new; p_new) moves into the product's epilogue.  r, p and Ap are double-buffered: every CTA reads set A
// anywhere and writes set B at its own rows, so one barrier per iteration orders everything; x is only touched at own rows.
// After the barrier every CTA sums the per-CTA partials of <r,r>, <Ap,Ap>, <p,Ap> (double-buffered by iteration parity) in the
// same fixed order and advances its own copy of the scalars (cg_advance, cg.hpp:170-180).  Per-entry arithmetic and the
// stopping rule are those of the two-kernel driver; only the grouping of the partial sums differs.
// ------------------------------------------------------------------------------------------------
// CTAS = resident CTAs per SM the kernel is compiled for: 2 (<= 128 registers: 24 gathers in flight per thread, 8 entries x 3 vectors,
// without spills) or 3 (<= 85 registers).
struct EpiCgOnePass
{
  const real *r0, *p0, *Ap0;     // previous iteration (read at any row)
  real *r1, *p1, *Ap1, *x;       // this iteration (written at own rows)
  real alpha, beta;
  real *partials;                // [3][VCL_MAX_BLOCKS]: <r,r>, <Ap,Ap>, <p,Ap>
  real acc[3];
  static constexpr int NQ = 3;
  static constexpr bool COO = false;
  static constexpr bool XFUSED = true;
  struct Pre { real r, p, Ap, x; };
  __device__ __forceinline__ real init(const Pre &) const { return 0.0; }
  __device__ __forceinline__ real term_scale() const { return 1.0; }
  __device__ __forceinline__ bool skip() const { return false; }
  __device__ __forceinline__ real xg(u32 c) const { return fma(beta, p0[c], fma(-alpha, Ap0[c], r0[c])); }
  __device__ __forceinline__ Pre pre(u32 i) const { Pre q; q.r = r0[i]; q.p = p0[i]; q.Ap = Ap0[i]; q.x = x[i]; return q; }
  __device__ __forceinline__ void row(u32 i, real dot, const Pre &q)
  {
    const real rn = fma(-alpha, q.Ap, q.r);                // the operations of cg_update_entries, in the same order
    const real pn = fma(beta, q.p, rn);
    x[i] = fma(alpha, q.p, q.x);
    r1[i] = rn; p1[i] = pn; Ap1[i] = dot;
    acc[0] = fma(rn, rn, acc[0]);
    acc[1] = fma(dot, dot, acc[1]);
    acc[2] = fma(pn, dot, acc[2]);
  }
  __device__ __forceinline__ void finish(real *smem)
  {
    block_sum<3>(acc, smem);
    if (threadIdx.x == 0)
    {
#pragma unroll
      for (int q = 0; q < 3; ++q) partials[q * VCL_MAX_BLOCKS + blockIdx.x] = acc[q];
    }
  }
};

// bufs: r, p, Ap of set A followed by set B; st->pad0 tells which set is current (0: A) and is updated on exit
template<int CTAS>
__global__ void __launch_bounds__(CSR_BLOCK_THREADS, CTAS)
cg_onepass_kernel(CsrDev A, real *x, real *rA, real *pA, real *ApA, real *rB, real *pB, real *ApB,
                  SolverState *st, real *partials, int iterations)
{
  cgrp::grid_group grid = cgrp::this_grid();
  __shared__ SolverState s_st;
  __shared__ real s_sum[3 * 32];
  if (threadIdx.x == 0) s_st = *st;
  __syncthreads();
  CsrCarry carry = {0u, 0};
  const XVec xv = {x, (u32)sizeof(real), nullptr, 0u};    // unit stride (the float build's fuse pattern looks at it); entries come from xg()
  int cur = s_st.pad0;

  for (int it = 0; it < iterations; ++it)
  {
    if (s_st.done != VCL_RUNNING) break;                   // identical in every CTA
    real *part = partials + (size_t)(it & 1) * 3 * VCL_MAX_BLOCKS;    // a CTA already in iteration it+1 must not overwrite what a slower one still sums
    {
      EpiCgOnePass epi = {cur ? rB : rA, cur ? pB : pA, cur ? ApB : ApA, cur ? rA : rB, cur ? pA : pB, cur ? ApA : ApB, x,
                          s_st.alpha, s_st.beta, part, {0.0, 0.0, 0.0}};
      csr_stream_body<EpiCgOnePass, false, true>(A, xv, epi, &carry);
    }
    grid.sync();                                           // the new r / p / Ap and all partials are visible
    real tot[3];
    sum_partials<3>(part, tot, s_sum);
    if (threadIdx.x == 0)
    {
      s_st.sums[0] = tot[0]; s_st.sums[1] = tot[1]; s_st.sums[2] = tot[2];
      cg_advance(&s_st);                                   // cg.hpp:170-180
    }
    __syncthreads();
    cur ^= 1;
  }
  {
    EpiCgOnePass epi = {rA, pA, ApA, rB, pB, ApB, x, 0.0, 0.0, partials, {0.0, 0.0, 0.0}};
    csr_stream_body<EpiCgOnePass, false, true>(A, xv, epi, &carry, true);        // wait for the copy issued ahead, if any
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { s_st.pad0 = cur; *st = s_st; }
}
// ------------------------------------------------------------------------------------------------
// Jacobi / row-scaling preconditioned CG (single-reduction form, see pcg_update_kernel) in the same persistent form:
//     p = u + beta p; s = w + beta s; x += alpha p; r -= alpha s; u = r ./ diag    -> partial gamma = <r,u>
//     grid barrier;   w = A u (csr_stream_body)                                      -> partial delta = <w,u>
//     grid barrier;   every CTA sums the partials and advances alpha / beta / the convergence test (pcg_advance)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CSR_BLOCK_THREADS, CSR_MIN_CTAS)
pcg_persistent_kernel(CsrDev A, XVec xv, long long n, real *x, real *r, real *u, real *w, real *p, real *s, const real *diag,
                      SolverState *st, real *partials, int iterations)
{
  cgrp::grid_group grid = cgrp::this_grid();
  __shared__ SolverState s_st;
  __shared__ real s_sum[2 * 32];
  if (threadIdx.x == 0) s_st = *st;
  __syncthreads();
  CsrCarry carry = {0u, 0};
  for (int it = 0; it < iterations; ++it)
  {
    if (s_st.done != VCL_RUNNING) break;
    real *part_gamma = partials + (2 + (it & 1)) * VCL_MAX_BLOCKS;     // double-buffered like <r,r> in cg_persistent_kernel
    {
      real acc[1] = {pcg_update_entries(n, x, r, u, w, p, s, diag, s_st.alpha, s_st.beta)};
      block_sum<1>(acc, s_sum);
      if (threadIdx.x == 0) part_gamma[blockIdx.x] = acc[0];
    }
    grid.sync();
    EpiCgPartial epi = {w, u, partials, {0.0, 0.0}};       // partial arrays [0]: <w,w> (unused), [1]: <u,w>
    csr_stream_body<EpiCgPartial, false, true>(A, xv, epi, &carry);
    grid.sync();
    real dg[2];
    const real * const arrs[2] = {partials + VCL_MAX_BLOCKS, part_gamma};
    sum_partials_of<2>(arrs, dg, s_sum);                   // delta = <w,u>, gamma = <r,u>
    if (threadIdx.x == 0)
    {
      s_st.sums[0] = dg[1];
      pcg_advance(&s_st, dg[0]);
    }
    __syncthreads();
  }
  {
    EpiCgPartial epi = {w, u, partials, {0.0, 0.0}};
    csr_stream_body<EpiCgPartial, false, true>(A, xv, epi, &carry, true);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *st = s_st;
}
// ------------------------------------------------------------------------------------------------
// Pipelined BiCGStab (bicgstab.hpp:97-215) in the persistent form.  Per iteration, four phases and four grid barriers:
//     Ap = A p                          -> partials <Ap,r0*>                    | barrier | alpha = <r,r0*> / <Ap,r0*>
//     s = r - alpha Ap                  -> partial  <s,s>                       | barrier |
//     As = A s                          -> partials <As,As>, <As,s>, <As,r0*>   | barrier | beta, omega, residual, convergence test
//     x += alpha p + omega s; r = s - omega As; p = r + beta (p - omega Ap)  -> partial <r,r0*> | barrier |
// (the stand-alone form needs four launches for the same work).  On convergence the loop leaves BEFORE the last vector
// update, like the reference (bicgstab.hpp:196-206).  Both products go through the same csr_stream_body instantiation, so
// the descriptors and the copy issued ahead are shared between them.
// ------------------------------------------------------------------------------------------------
struct EpiDot3Partial
{
  real *out; const real *in; const real *r0; real *partials;
  real acc[3];
  static constexpr int NQ = 3;
  static constexpr bool COO = false;
  struct Pre { real v, r0; };
  __device__ __forceinline__ real init(const Pre &) const { return 0.0; }
  __device__ __forceinline__ real term_scale() const { return 1.0; }
  __device__ __forceinline__ bool skip() const { return false; }
  __device__ __forceinline__ Pre pre(u32 r) const { Pre q; q.v = in[r]; q.r0 = r0[r]; return q; }
  __device__ __forceinline__ void row(u32 r, real dot, const Pre &q)
  {
    out[r] = dot;
    acc[0] = fma(dot, dot, acc[0]);
    acc[1] = fma(q.v, dot, acc[1]);
    acc[2] = fma(dot, q.r0, acc[2]);
  }
  __device__ __forceinline__ void finish(real *smem)
  {
    block_sum<3>(acc, smem);
    if (threadIdx.x == 0)
    {
#pragma unroll
      for (int q = 0; q < 3; ++q) partials[q * VCL_MAX_BLOCKS + blockIdx.x] = acc[q];
    }
  }
};

__global__ void __launch_bounds__(CSR_BLOCK_THREADS, CSR_MIN_CTAS)
bicgstab_persistent_kernel(CsrDev A, long long n, real *x, real *r, real *p, const real *r0, real *Ap, real *s, real *As,
                           SolverState *st, real *partials, int iterations)
{
  cgrp::grid_group grid = cgrp::this_grid();
  __shared__ SolverState s_st;
  __shared__ real s_sum[4 * 32];
  if (threadIdx.x == 0) s_st = *st;
  __syncthreads();
  CsrCarry carry = {0u, 0};
  real *part_ss = partials + 3 * VCL_MAX_BLOCKS, *part_rr0 = partials + 4 * VCL_MAX_BLOCKS;
  const XVec xv_p = {p, (u32)sizeof(real), nullptr, 0u}, xv_s = {s, (u32)sizeof(real), nullptr, 0u};

  for (int it = 0; it < iterations; ++it)
  {
    if (s_st.done != VCL_RUNNING) break;
    // ---- Ap = A p ----
    {
      EpiDot3Partial e = {Ap, p, r0, partials, {0.0, 0.0, 0.0}};
      csr_stream_body<EpiDot3Partial, false, true>(A, xv_p, e, &carry);
    }
    grid.sync();
    {
      real t[1];
      sum_partials<1>(partials + 2 * VCL_MAX_BLOCKS, t, s_sum);
      if (threadIdx.x == 0) s_st.sums[3] = t[0];                                   // <Ap,r0*>
      __syncthreads();
    }
    // ---- s = r - alpha Ap, <s,s> ----
    {
      real acc[1] = {bicgstab_s_entries(n, s, r, Ap, s_st.sums[0] / s_st.sums[3])};
      block_sum<1>(acc, s_sum);
      if (threadIdx.x == 0) part_ss[blockIdx.x] = acc[0];
    }
    grid.sync();
    // ---- As = A s ----
    {
      EpiDot3Partial e = {As, s, r0, partials, {0.0, 0.0, 0.0}};
      csr_stream_body<EpiDot3Partial, false, true>(A, xv_s, e, &carry);
    }
    grid.sync();
    {
      real t[4];
      const real * const arrs[4] = {partials, partials + VCL_MAX_BLOCKS, partials + 2 * VCL_MAX_BLOCKS, part_ss};
      sum_partials_of<4>(arrs, t, s_sum);
      if (threadIdx.x == 0)
      {
        s_st.sums[1] = t[0]; s_st.sums[2] = t[1]; s_st.sums[4] = t[2]; s_st.sums[5] = t[3];
        bicgstab_advance(&s_st);                                                   // bicgstab.hpp:184-199
      }
      __syncthreads();
    }
    if (s_st.done != VCL_RUNNING) break;                                           // converged: the iterate before this update is returned
    // ---- x += alpha p + omega s; r = s - omega As; p = r + beta (p - omega Ap); <r,r0*> ----
    {
      real acc[1] = {bicgstab_update_entries(n, x, p, s, r, As, Ap, r0, s_st.alpha, s_st.beta, s_st.omega)};
      block_sum<1>(acc, s_sum);
      if (threadIdx.x == 0) part_rr0[blockIdx.x] = acc[0];
    }
    grid.sync();
    {
      real t[1];
      sum_partials<1>(part_rr0, t, s_sum);
      if (threadIdx.x == 0)
      {
        s_st.sums[0] = t[0];                                                       // <r,r0*>
        if (s_st.iters >= s_st.maxit) s_st.done = VCL_MAXIT;
      }
      __syncthreads();
    }
  }
  {
    EpiDot3Partial e = {As, s, r0, partials, {0.0, 0.0, 0.0}};
    csr_stream_body<EpiDot3Partial, false, true>(A, xv_s, e, &carry, true);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *st = s_st;
}
// ------------------------------------------------------------------------------------------------
// One restart cycle of the pipelined GMRES (gmres.hpp:241-300: product, classical Gram-Schmidt in two stages, normalisation)
// in the persistent form.  Inner iteration k (v_k = column k of the basis, src = the residual for k = 0, else v_{k-1}):
//     v_k = A src [./ diag]                         -> partial ||v_k||^2 (used for k = 0)                  | barrier
//     k > 0: partial h_j = <v_j, v_k>, j < k        (4 columns per pass over this CTA's row share)          | barrier
//            CTA c sums the partials of column c    -> h[c], R[c + k m]                                    | barrier
//            v_k -= sum_j h_j v_j                   -> partial ||v_k||^2                                   | barrier
//     R[k + k m] = ||v_k||; v_k /= ||v_k||          -> partial xi_k = <res, v_k> (summed by CTA 0)         | barrier
// The stand-alone form launches four kernels per inner iteration.  Scratch arrays of the backend: partials[j] for the
// h_j (j <= 59), partials[60] for ||v_k||^2, partials[61] for xi_k -- hence krylov_dim <= 60 on this path.
// ------------------------------------------------------------------------------------------------
#define GMRES_PERSISTENT_MAX_KRYLOV 60
struct EpiNsqPartial
{
  real *out; const real *diag; real *partials;
  real acc[1];
  static constexpr int NQ = 1;
  static constexpr bool COO = false;
  typedef real Pre;
  __device__ __forceinline__ real init(real) const { return 0.0; }
  __device__ __forceinline__ real term_scale() const { return 1.0; }
  __device__ __forceinline__ bool skip() const { return false; }
  __device__ __forceinline__ real pre(u32 r) const { return diag ? diag[r] : real(1); }
  __device__ __forceinline__ void row(u32 r, real dot, real d)
  {
    if (diag) dot = dot / d;
    out[r] = dot;
    acc[0] = fma(dot, dot, acc[0]);
  }
  __device__ __forceinline__ void finish(real *smem)
  {
    block_sum<1>(acc, smem);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc[0];
  }
};

__global__ void __launch_bounds__(CSR_BLOCK_THREADS, CSR_MIN_CTAS)
gmres_persistent_kernel(CsrDev A, long long n, long long isz, int m, const real *res, real *V, const real *diag,
                        real *R, real *d_h, real *d_xi, real *partials)
{
  cgrp::grid_group grid = cgrp::this_grid();
  __shared__ real s_sum[8 * 32];
  __shared__ real s_h[VCL_GMRES_MAX_KRYLOV];
  __shared__ real s_nsq;
  CsrCarry carry = {0u, 0};
  real *part_nsq = partials + 60 * VCL_MAX_BLOCKS, *part_xi = partials + 61 * VCL_MAX_BLOCKS;
  const long long tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  const long long npairs = n >> 1;                         // the basis is 16-byte aligned with an even internal size (checked by the host)
  const bool res_vec = (reinterpret_cast<uintptr_t>(res) & 15u) == 0u;

  for (int k = 0; k < m; ++k)
  {
    real *vk = V + (size_t)k * isz;
    const real *src = (k == 0) ? res : V + (size_t)(k - 1) * isz;
    {
      const XVec xv = {src, (u32)sizeof(real), nullptr, 0u};
      EpiNsqPartial e = {vk, diag, part_nsq, {0.0}};
      csr_stream_body<EpiNsqPartial, false, true>(A, xv, e, &carry);
    }
    grid.sync();
    if (k > 0)
    {
      // ---- stage 1: partial h_j = <v_j, v_k> over this CTA's rows, GS columns per pass (4: the 64-register budget of 4 CTAs/SM) ----
      constexpr int GS = 4;
      for (int c0 = 0; c0 < k; c0 += GS)
      {
        const int nc = min(GS, k - c0);
        const real *col = V + (size_t)c0 * isz;
        real acc[GS];
#pragma unroll
        for (int q = 0; q < GS; ++q) acc[q] = 0.0;
        for (long long pi = tid0; pi < npairs; pi += nthr)
        {
          const real2 v = ld2(vk, 2 * pi);
          real2 a[GS];
#pragma unroll
          for (int q = 0; q < GS; ++q)
            if (q < nc) a[q] = ld2(col + (size_t)q * isz, 2 * pi);
#pragma unroll
          for (int q = 0; q < GS; ++q)
            if (q < nc) { acc[q] = fma(a[q].x, v.x, acc[q]); acc[q] = fma(a[q].y, v.y, acc[q]); }
        }
        if ((n & 1) && tid0 == 0)
        {
          const real v = vk[n - 1];
#pragma unroll
          for (int q = 0; q < GS; ++q)
            if (q < nc) acc[q] = fma(col[(size_t)q * isz + n - 1], v, acc[q]);
        }
        block_sum<GS>(acc, s_sum);
        if (threadIdx.x == 0)
        {
#pragma unroll
          for (int q = 0; q < GS; ++q)
            if (q < nc) partials[(size_t)(c0 + q) * VCL_MAX_BLOCKS + blockIdx.x] = acc[q];
        }
      }
      grid.sync();
      // ---- column c is summed by CTA c (c, c + grid, ...) ----
      for (int c = blockIdx.x; c < k; c += gridDim.x)
      {
        real t[1];
        sum_partials<1>(partials + (size_t)c * VCL_MAX_BLOCKS, t, s_sum);
        if (threadIdx.x == 0) { d_h[c] = t[0]; R[(size_t)c + (size_t)k * m] = t[0]; }
      }
      grid.sync();
      for (int j = threadIdx.x; j < k; j += blockDim.x) s_h[j] = __ldcg(d_h + j);
      __syncthreads();
      // ---- stage 2: v_k -= sum_j h_j v_j, partial ||v_k||^2 ----
      {
        real acc[1] = {0.0};
        for (long long pi = tid0; pi < npairs; pi += nthr)
        {
          real2 v = ld2(vk, 2 * pi);
          int j = 0;
          for (; j + 4 <= k; j += 4)
          {
            const real2 a0 = ld2(V + (size_t)j * isz, 2 * pi), a1 = ld2(V + (size_t)(j + 1) * isz, 2 * pi);
            const real2 a2 = ld2(V + (size_t)(j + 2) * isz, 2 * pi), a3 = ld2(V + (size_t)(j + 3) * isz, 2 * pi);
            v.x = fma(-s_h[j], a0.x, v.x);     v.y = fma(-s_h[j], a0.y, v.y);
            v.x = fma(-s_h[j + 1], a1.x, v.x); v.y = fma(-s_h[j + 1], a1.y, v.y);
            v.x = fma(-s_h[j + 2], a2.x, v.x); v.y = fma(-s_h[j + 2], a2.y, v.y);
            v.x = fma(-s_h[j + 3], a3.x, v.x); v.y = fma(-s_h[j + 3], a3.y, v.y);
          }
          for (; j < k; ++j)
          {
            const real2 a0 = ld2(V + (size_t)j * isz, 2 * pi);
            v.x = fma(-s_h[j], a0.x, v.x); v.y = fma(-s_h[j], a0.y, v.y);
          }
          acc[0] = fma(v.x, v.x, acc[0]); acc[0] = fma(v.y, v.y, acc[0]);
          st2(vk, 2 * pi, v);
        }
        if ((n & 1) && tid0 == 0)
        {
          real v = vk[n - 1];
          for (int j = 0; j < k; ++j) v = fma(-s_h[j], V[(size_t)j * isz + n - 1], v);
          acc[0] = fma(v, v, acc[0]);
          vk[n - 1] = v;
        }
        block_sum<1>(acc, s_sum);
        if (threadIdx.x == 0) part_nsq[blockIdx.x] = acc[0];
      }
      grid.sync();
    }
    // ---- ||v_k|| (every CTA), normalisation, partial xi_k = <res, v_k> ----
    {
      real t[1];
      sum_partials<1>(part_nsq, t, s_sum);
      if (threadIdx.x == 0) s_nsq = t[0];
      __syncthreads();
      const real nrm = sqrt(s_nsq);
      if (tid0 == 0) R[(size_t)k * m + k] = nrm;
      real acc[1] = {0.0};
      const long long np2 = res_vec ? npairs : 0;
      for (long long pi = tid0; pi < np2; pi += nthr)
      {
        real2 v = ld2(vk, 2 * pi); const real2 rr = ld2(res, 2 * pi);
        v.x = v.x / nrm; v.y = v.y / nrm;
        acc[0] = fma(rr.x, v.x, acc[0]); acc[0] = fma(rr.y, v.y, acc[0]);
        st2(vk, 2 * pi, v);
      }
      for (long long i = 2 * np2 + tid0; i < n; i += nthr)
      {
        const real v = vk[i] / nrm;
        acc[0] = fma(res[i], v, acc[0]);
        vk[i] = v;
      }
      block_sum<1>(acc, s_sum);
      if (threadIdx.x == 0) part_xi[blockIdx.x] = acc[0];
    }
    grid.sync();
    if (blockIdx.x == 0)
    {
      real t[1];
      sum_partials<1>(part_xi, t, s_sum);
      if (threadIdx.x == 0) d_xi[k] = t[0];
    }
  }
  {
    const XVec xv = {res, (u32)sizeof(real), nullptr, 0u};
    EpiNsqPartial e = {V, diag, part_nsq, {0.0}};
    csr_stream_body<EpiNsqPartial, false, true>(A, xv, e, &carry, true);
  }
}
}
