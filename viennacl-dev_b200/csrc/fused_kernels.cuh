// fused_kernels.cuh -- fused vector-update + inner-product kernels of the pipelined Krylov solvers and the
// fused SpMV epilogues.  Replaces cuda/iterative_operations.hpp:43-103 (K5), :733-886 (K9, K10), :1597-1894 (K14-K17)
// and the reduction halves of K6-K8 / K11-K13.
//
// Common design: persistent grid (a few CTAs per SM), 16-byte vector loads, per-CTA partial sums reduced in a fixed order
// by whichever CTA finishes last (grid_sum_last_block) -> results are run-to-run deterministic; the last CTA also advances
// the solver's scalar recurrences in device memory (SolverState), which removes the per-iteration blocking D2H copy of the
// reference drivers (cg.hpp:168, bicgstab.hpp:180).
#pragma once
#include "common.cuh"
#include "prec.cuh"
#include "solver_state.cuh"
#include "peer.cuh"

namespace VCL_NS
{

#define VEC_THREADS 256

// ------------------------------------------------------------------------------------------------
// Scalar recurrences (run by one thread)
// ------------------------------------------------------------------------------------------------
// cg.hpp:170-180.  sums[0] = <r,r>, sums[1] = <Ap,Ap>, sums[2] = <p,Ap>
__device__ __forceinline__ void cg_advance(SolverState *st)
{
  const real rr = st->sums[0], ApAp = st->sums[1], pAp = st->sums[2];
  st->iters += 1;
  st->est = sqrt(fabs(rr / st->norm_rhs_sq));
  if (fabs(rr / st->norm_rhs_sq) < st->tol * st->tol || fabs(rr) < st->abs_tol * st->abs_tol) { st->done = VCL_CONVERGED; return; }
  if (st->iters >= st->maxit) { st->done = VCL_MAXIT; return; }
  const real alpha = rr / pAp;
  st->alpha = alpha;
  st->beta = (alpha * alpha * ApAp - rr) / rr;
}

// Jacobi-preconditioned CG in the single-reduction form of Chronopoulos/Gear (the recurrence the reference's pipelined CG
// is built on, cg.hpp:116-118, extended by the preconditioner; same iterates as the classical PCG of cg.hpp:257-322 up to
// rounding).  sums[0] = gamma = <r, u> with u = r ./ diag (written by pcg_update_kernel), delta = <w, u> with w = A u.
// Bookkeeping as in cg.hpp:296-309: iteration counted, estimate sqrt(|gamma / gamma_0|), squared tolerances.
__device__ __forceinline__ void pcg_advance(SolverState *st, real delta)
{
  const real gamma = st->sums[0];
  st->iters += 1;
  st->est = sqrt(fabs(gamma / st->norm_rhs_sq));
  if (fabs(gamma / st->norm_rhs_sq) < st->tol * st->tol || fabs(gamma) < st->abs_tol * st->abs_tol) { st->done = VCL_CONVERGED; return; }
  if (st->iters >= st->maxit) { st->done = VCL_MAXIT; return; }
  const real beta = gamma / st->ip_rr0;                          // ip_rr0 holds the previous gamma
  st->alpha = gamma / (delta - beta * gamma / st->alpha);
  st->beta = beta;
  st->ip_rr0 = gamma;
}

// bicgstab.hpp:184-199.  chunks: 0 <r,r0*>, 1 <As,As>, 2 <As,s>, 3 <Ap,r0*>, 4 <As,r0*>, 5 <s,s>
__device__ __forceinline__ void bicgstab_advance(SolverState *st)
{
  const real r_r0 = st->sums[0], As_As = st->sums[1], As_s = st->sums[2], Ap_r0 = st->sums[3], As_r0 = st->sums[4], s_s = st->sums[5];
  st->iters += 1;
  st->alpha = r_r0 / Ap_r0;
  st->beta = -As_r0 / Ap_r0;
  const real omega = As_s / As_As;
  st->omega = omega;
  const real res = sqrt(s_s - 2.0 * omega * As_s + omega * omega * As_As);
  st->residual_norm = res;
  st->est = fabs(res / st->norm_rhs);
  if (fabs(res / st->norm_rhs) < st->tol || res < st->abs_tol) st->done = VCL_CONVERGED;
}

// ------------------------------------------------------------------------------------------------
// Fused SpMV epilogue:  Ap[r] = dot (optionally / diag[r]);  <Ap,Ap>, <p,Ap>, <Ap,r0*>
// host_based/iterative_operations.hpp:58-103.  STEP selects what the last CTA does after the reduction.
// ------------------------------------------------------------------------------------------------
enum { DIST_CG = 1, DIST_BICG_P = 2, DIST_BICG_S = 3, DIST_PCG = 4 };
enum { STEP_NONE = 0, STEP_CG = 1, STEP_BICGSTAB = 2, STEP_PBICG_ALPHA = 3, STEP_PBICG_OMEGA = 4, STEP_PCG = 5 };

template<int STEP, bool USE_R0, bool JACOBI>
struct EpiFused
{
  real *Ap; const real *p; const real *r0; const real *diag;
  real *partials; unsigned int *ticket;
  SolverState *st;                 // NULL in per-op API mode
  real *out0, *out1, *out2;      // totals: <Ap,Ap>, <p,Ap>, <Ap,r0*>
  real acc[3];
  const real *add_from;          // optional: 3 totals of an earlier launch over a disjoint row subset (interior + boundary split)
  // row-partitioned solvers over peer memory (peer.cuh): the last CTA all-reduces this launch's rank-local totals together
  // with ONE rank-local total of the preceding vector kernel (*loc_extra) across the ranks and then advances the solver's
  // scalars -- identical on every rank.  win == NULL: single-domain behaviour.  dist_mode (what is reduced -> where it goes):
  //   DIST_CG       {*loc <r,r>, <Ap,Ap>, <p,Ap>}                  -> sums[0..2], cg_advance
  //   DIST_BICG_P   {*loc <r,r0*>, <Ap,r0*>}                        -> sums[0], sums[3]                 (Ap = A p)
  //   DIST_BICG_S   {*loc <s,s>, <As,As>, <As,s>, <As,r0*>}         -> sums[5], sums[1], sums[2], sums[4], bicgstab_advance
  //   DIST_PCG      {*loc gamma = <r,u>, delta = <w,u>}             -> sums[0], pcg_advance(delta)
  const PeerWindow *win; unsigned long long red_seq; const real *loc_extra; int dist_mode;
  static constexpr int NQ = 3;
  static constexpr bool COO = false;
  // the epilogue's own per-row operands, all requested BEFORE the row's gather chain so that their latency overlaps with it
  struct Pre { real p, d, r0; };
  __device__ __forceinline__ real init(const Pre &) const { return 0.0; }
  __device__ __forceinline__ real term_scale() const { return 1.0; }

  __device__ __forceinline__ bool skip() const { return st != nullptr && (st->done != VCL_RUNNING || st->need_restart != 0); }
  __device__ __forceinline__ Pre pre(u32 r) const
  {
    Pre q;
    q.p = p[r];
    q.d = JACOBI ? diag[r] : real(1);
    q.r0 = USE_R0 ? r0[r] : real(0);
    return q;
  }
  __device__ __forceinline__ void row(u32 r, real dot, const Pre &q)
  {
    if (JACOBI) dot = dot / q.d;
    Ap[r] = dot;
    acc[0] = fma(dot, dot, acc[0]);
    acc[1] = fma(q.p, dot, acc[1]);
    if (USE_R0) acc[2] = fma(dot, q.r0, acc[2]);
  }
  __device__ __forceinline__ void finish(real *smem)
  {
    if (!grid_sum_last_block<3>(acc, partials, ticket, smem)) return;
#ifndef VCL_F32          // the row-partitioned path (peer.cuh) exists in double precision only
    if (win != nullptr)
    {
#ifdef VCL_PEER_DEBUG
      const u64 t_a = global_ns();
#endif
      if (threadIdx.x == 0)
      {
        smem[0] = *loc_extra; smem[1] = 0.0; smem[2] = 0.0; smem[3] = 0.0;
        if (dist_mode == DIST_CG)          { smem[1] = acc[0]; smem[2] = acc[1]; }
        else if (dist_mode == DIST_BICG_P) { smem[1] = acc[2]; }
        else if (dist_mode == DIST_BICG_S) { smem[1] = acc[0]; smem[2] = acc[1]; smem[3] = acc[2]; }
        else                               { smem[1] = acc[1]; }
      }
      peer_allreduce<4>(win, red_seq, smem, smem + 32);
      if (threadIdx.x == 0)
      {
        if (dist_mode == DIST_CG)          { st->sums[0] = smem[0]; st->sums[1] = smem[1]; st->sums[2] = smem[2]; cg_advance(st); }
        else if (dist_mode == DIST_BICG_P) { st->sums[0] = smem[0]; st->sums[3] = smem[1]; }
        else if (dist_mode == DIST_BICG_S) { st->sums[5] = smem[0]; st->sums[1] = smem[1]; st->sums[2] = smem[2]; st->sums[4] = smem[3]; bicgstab_advance(st); }
        else                               { st->sums[0] = smem[0]; pcg_advance(st, smem[1]); }
#ifdef VCL_PEER_DEBUG
        if (win->dbg) { u64 *d = win->dbg + (red_seq % 1024) * 4; d[0] = t_a; d[1] = global_ns(); }
#endif
      }
      return;
    }
#endif
    if (threadIdx.x == 0)
    {
      if (add_from) { acc[0] += add_from[0]; acc[1] += add_from[1]; if (USE_R0) acc[2] += add_from[2]; }
      if (out0) *out0 = acc[0];
      if (out1) *out1 = acc[1];
      if (USE_R0 && out2) *out2 = acc[2];
      if (STEP == STEP_CG) cg_advance(st);
      if (STEP == STEP_BICGSTAB) bicgstab_advance(st);
      if (STEP == STEP_PCG) pcg_advance(st, acc[1]);
      if (STEP == STEP_PBICG_ALPHA) st->alpha = st->ip_rr0 / acc[2];                       // bicgstab.hpp:449
      if (STEP == STEP_PBICG_OMEGA) { const real nt = sqrt(acc[0]); st->omega = acc[1] / (nt * nt); }  // bicgstab.hpp:455-456
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Helpers for streaming vector kernels: every thread handles pairs (16-byte accesses); a scalar tail covers odd n.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool aligned16(const void *a, const void *b = nullptr, const void *c = nullptr, const void *d = nullptr,
                                          const void *e = nullptr, const void *f = nullptr, const void *g = nullptr)
{
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(d) |
           reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(g)) & 15u) == 0u;
}
__device__ __forceinline__ real2 ld2(const real *p, long long i) { return *reinterpret_cast<const real2*>(p + i); }
__device__ __forceinline__ void st2(real *p, long long i, real2 v) { *reinterpret_cast<real2*>(p + i) = v; }

// ------------------------------------------------------------------------------------------------
// CG: x += alpha p; r -= alpha Ap; p = r + beta p; <r,r>        (host_based/iterative_operations.hpp:378-418)
// ------------------------------------------------------------------------------------------------
// the entries of this thread: returns its share of <r,r>.  Shared by the stand-alone kernel and the persistent one.
template<bool PUSH>
__device__ __forceinline__ real cg_update_entries(long long n, real *x, real *p, real *r, const real *Ap, real alpha, real beta, const PushRanges &pr,
                                                  bool *pushed = nullptr)
{
  real acc = 0.0;
  bool sent = false;
  const long long npairs = aligned16(x, p, r, Ap) ? (n >> 1) : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += (long long)gridDim.x * blockDim.x)
  {
    const long long k = i * 2;
    real2 vx = ld2(x, k), vp = ld2(p, k), vr = ld2(r, k); const real2 va = ld2(Ap, k);
    vx.x = fma(alpha, vp.x, vx.x);       vx.y = fma(alpha, vp.y, vx.y);
    vr.x = fma(-alpha, va.x, vr.x);      vr.y = fma(-alpha, va.y, vr.y);
    vp.x = fma(beta, vp.x, vr.x);        vp.y = fma(beta, vp.y, vr.y);
    acc = fma(vr.x, vr.x, acc);          acc = fma(vr.y, vr.y, acc);
    st2(x, k, vx); st2(r, k, vr); st2(p, k, vp);
    if (PUSH && pr.n) { sent |= push_entry(pr, k, vp.x); sent |= push_entry(pr, k + 1, vp.y); }      // new p -> the neighbours' halo buffers (NVLink)
  }
  for (long long k = 2 * npairs + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
  {
    real vp = p[k], vr = r[k];
    x[k] = fma(alpha, vp, x[k]);
    vr = fma(-alpha, Ap[k], vr);
    vp = fma(beta, vp, vr);
    acc = fma(vr, vr, acc);
    p[k] = vp; r[k] = vr;
    if (PUSH && pr.n) sent |= push_entry(pr, k, vp);
  }
  if (pushed) *pushed = sent;
  return acc;
}

template<bool PUSH>
__device__ __forceinline__ void cg_update_body(long long n, real *x, real *p, real *r, const real *Ap, real alpha_v, real beta_v,
                                               SolverState *st, real *partials, unsigned int *ticket, real *out_rr, const PushRanges &pr)
{
  __shared__ real s_red[32];
  if (st != nullptr && st->done != VCL_RUNNING) return;
  const real alpha = st ? st->alpha : alpha_v;
  const real beta  = st ? st->beta  : beta_v;
  bool pushed = false;
  real acc[1] = {cg_update_entries<PUSH>(n, x, p, r, Ap, alpha, beta, pr, &pushed)};
  if (pushed) __threadfence_system();               // THIS thread's remote stores (3 % of the threads have any) are performed before its
                                                    // CTA takes a ticket: grid_sum_last_block synchronises the CTA first
  if (grid_sum_last_block<1>(acc, partials, ticket, s_red))
  {
    if (threadIdx.x == 0) *out_rr = acc[0];
    // every CTA has fenced its pushes: publish the sequence number to the destinations
    if (PUSH && (int)threadIdx.x < pr.n) { __threadfence_system(); st_release_sys(pr.flag[threadIdx.x], pr.seq); }
  }
}

// Single-domain form.  4 resident CTAs per SM (62 registers, no spills): compiled for 5 / 6 / 8 CTAs the kernel spills (8 / 80 / 144
// bytes of stack) and a 256^3 CG iteration goes from 416 to 414 / 452 / 485 us (profiles/ab_cgupdate_r2.log) -- not worth it.
static __global__ void __launch_bounds__(VEC_THREADS, 4)
cg_update_kernel(long long n, real *x, real *p, real *r, const real *Ap, real alpha_v, real beta_v,
                 SolverState *st, real *partials, unsigned int *ticket, real *out_rr)
{
  cg_update_body<false>(n, x, p, r, Ap, alpha_v, beta_v, st, partials, ticket, out_rr, PushRanges());
}

// Row-partitioned form: additionally writes the new p entries the neighbours need into their halo buffers (peer.cuh)
static __global__ void __launch_bounds__(VEC_THREADS)
cg_update_push_kernel(long long n, real *x, real *p, real *r, const real *Ap, real alpha_v, real beta_v,
                      SolverState *st, real *partials, unsigned int *ticket, real *out_rr, const PushRanges pr)
{
  cg_update_body<true>(n, x, p, r, Ap, alpha_v, beta_v, st, partials, ticket, out_rr, pr);
}

// ------------------------------------------------------------------------------------------------
// Jacobi-PCG update: p = u + beta p; s = w + beta s; x += alpha p; r -= alpha s; u = r ./ diag; <r,u>
// (one pass: 7 reads + 5 writes per entry; the reference's generic PCG makes ~10 passes and 2 blocking reductions)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ real pcg_update_entries(long long n, real *x, real *r, real *u, const real *w, real *p, real *s, const real *diag,
                                                   real alpha, real beta)
{
  real acc = 0.0;
  const long long npairs = aligned16(x, r, u, w, p, s, diag) ? (n >> 1) : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += (long long)gridDim.x * blockDim.x)
  {
    const long long k = i * 2;
    real2 vx = ld2(x, k), vr = ld2(r, k), vu = ld2(u, k), vp = ld2(p, k), vs = ld2(s, k);
    const real2 vw = ld2(w, k), vd = ld2(diag, k);
    vp.x = fma(beta, vp.x, vu.x);        vp.y = fma(beta, vp.y, vu.y);
    vs.x = fma(beta, vs.x, vw.x);        vs.y = fma(beta, vs.y, vw.y);
    vx.x = fma(alpha, vp.x, vx.x);       vx.y = fma(alpha, vp.y, vx.y);
    vr.x = fma(-alpha, vs.x, vr.x);      vr.y = fma(-alpha, vs.y, vr.y);
    vu.x = vr.x / vd.x;                  vu.y = vr.y / vd.y;
    acc = fma(vr.x, vu.x, acc);          acc = fma(vr.y, vu.y, acc);
    st2(p, k, vp); st2(s, k, vs); st2(x, k, vx); st2(r, k, vr); st2(u, k, vu);
  }
  for (long long k = 2 * npairs + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
  {
    const real vp = fma(beta, p[k], u[k]), vs = fma(beta, s[k], w[k]);
    x[k] = fma(alpha, vp, x[k]);
    const real vr = fma(-alpha, vs, r[k]);
    const real vu = vr / diag[k];
    acc = fma(vr, vu, acc);
    p[k] = vp; s[k] = vs; r[k] = vr; u[k] = vu;
  }
  return acc;
}

static __global__ void __launch_bounds__(VEC_THREADS)
pcg_update_kernel(long long n, real *x, real *r, real *u, const real *w, real *p, real *s, const real *diag,
                  SolverState *st, real *partials, unsigned int *ticket, real *out_gamma)
{
  __shared__ real s_red[32];
  if (st->done != VCL_RUNNING) return;
  real acc[1] = {pcg_update_entries(n, x, r, u, w, p, s, diag, st->alpha, st->beta)};
  if (grid_sum_last_block<1>(acc, partials, ticket, s_red) && threadIdx.x == 0) *out_gamma = acc[0];
}

// u = r ./ diag; <r,u>   (set-up of the Jacobi-PCG)
static __global__ void __launch_bounds__(VEC_THREADS)
pcg_init_kernel(long long n, const real *r, real *u, const real *diag, real *partials, unsigned int *ticket, real *out_gamma)
{
  __shared__ real s_red[32];
  real acc[1] = {0.0};
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
  {
    const real vu = r[k] / diag[k];
    acc[0] = fma(r[k], vu, acc[0]);
    u[k] = vu;
  }
  if (grid_sum_last_block<1>(acc, partials, ticket, s_red) && threadIdx.x == 0) *out_gamma = acc[0];
}

// ------------------------------------------------------------------------------------------------
// BiCGStab: s = r - alpha Ap with alpha = <r,r0*>/<Ap,r0*> taken from device memory; <s,s>
// (host_based/iterative_operations.hpp:518-563; cuda K9 :733-788 recomputes alpha in every CTA, here it is two loads)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ real bicgstab_s_entries(long long n, real *s, const real *r, const real *Ap, real alpha)
{
  real acc = 0.0;
  const long long npairs = aligned16(s, r, Ap) ? (n >> 1) : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += (long long)gridDim.x * blockDim.x)
  {
    const long long k = i * 2;
    const real2 vr = ld2(r, k), va = ld2(Ap, k);
    real2 vs;
    vs.x = fma(-alpha, va.x, vr.x); vs.y = fma(-alpha, va.y, vr.y);
    acc = fma(vs.x, vs.x, acc); acc = fma(vs.y, vs.y, acc);
    st2(s, k, vs);
  }
  for (long long k = 2 * npairs + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
  {
    const real vs = fma(-alpha, Ap[k], r[k]);
    acc = fma(vs, vs, acc);
    s[k] = vs;
  }
  return acc;
}

static __global__ void __launch_bounds__(VEC_THREADS)
bicgstab_update_s_kernel(long long n, real *s, const real *r, const real *Ap,
                         const real *in_r_r0, const real *in_Ap_r0,
                         SolverState *st, real *partials, unsigned int *ticket, real *out_ss)
{
  __shared__ real s_red[32];
  if (st != nullptr && st->done != VCL_RUNNING) return;
  real acc[1] = {bicgstab_s_entries(n, s, r, Ap, (*in_r_r0) / (*in_Ap_r0))};
  if (grid_sum_last_block<1>(acc, partials, ticket, s_red) && threadIdx.x == 0) *out_ss = acc[0];
}

// x += alpha p + omega s;  r = s - omega As;  p = r + beta (p - omega Ap);  <r,r0*>     (host_based/iterative_operations.hpp:572-621)
__device__ __forceinline__ real bicgstab_update_entries(long long n, real *x, real *p, const real *s, real *r, const real *As, const real *Ap,
                                                        const real *r0, real alpha, real beta, real omega)
{
  real acc = 0.0;
  const long long npairs = aligned16(x, p, s, r, As, Ap, r0) ? (n >> 1) : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += (long long)gridDim.x * blockDim.x)
  {
    const long long k = i * 2;
    real2 vx = ld2(x, k), vp = ld2(p, k); const real2 vs = ld2(s, k), vAs = ld2(As, k), vAp = ld2(Ap, k), v0 = ld2(r0, k);
    real2 vr;
    vx.x += alpha * vp.x + omega * vs.x;             vx.y += alpha * vp.y + omega * vs.y;
    vr.x = fma(-omega, vAs.x, vs.x);                 vr.y = fma(-omega, vAs.y, vs.y);
    vp.x = fma(beta, fma(-omega, vAp.x, vp.x), vr.x); vp.y = fma(beta, fma(-omega, vAp.y, vp.y), vr.y);
    acc = fma(vr.x, v0.x, acc);                      acc = fma(vr.y, v0.y, acc);
    st2(x, k, vx); st2(r, k, vr); st2(p, k, vp);
  }
  for (long long k = 2 * npairs + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
  {
    real vp = p[k]; const real vs = s[k];
    x[k] += alpha * vp + omega * vs;
    const real vr = fma(-omega, As[k], vs);
    vp = fma(beta, fma(-omega, Ap[k], vp), vr);
    acc = fma(vr, r0[k], acc);
    r[k] = vr; p[k] = vp;
  }
  return acc;
}

static __global__ void __launch_bounds__(VEC_THREADS)
bicgstab_update_kernel(long long n, real *x, real alpha_v, real *p, real omega_v, const real *s,
                       real *r, const real *As, real beta_v, const real *Ap, const real *r0,
                       SolverState *st, real *partials, unsigned int *ticket, real *out_r_r0)
{
  __shared__ real s_red[32];
  if (st != nullptr && st->done != VCL_RUNNING) return;
  const real alpha = st ? st->alpha : alpha_v;
  const real beta  = st ? st->beta  : beta_v;
  const real omega = st ? st->omega : omega_v;
  real acc[1] = {bicgstab_update_entries(n, x, p, s, r, As, Ap, r0, alpha, beta, omega)};
  if (grid_sum_last_block<1>(acc, partials, ticket, s_red) && threadIdx.x == 0)
  {
    *out_r_r0 = acc[0];
    if (st != nullptr && st->iters >= st->maxit) st->done = VCL_MAXIT;
  }
}

// ------------------------------------------------------------------------------------------------
// Left-preconditioned BiCGStab (bicgstab.hpp:398-489), device-resident scalars
// ------------------------------------------------------------------------------------------------
// s = r - alpha t0                                            (bicgstab.hpp:451)
static __global__ void __launch_bounds__(VEC_THREADS)
pbicg_s_kernel(long long n, real *s, const real *r, const real *t0, const SolverState *st)
{
  if (st->done != VCL_RUNNING || st->need_restart) return;
  const real alpha = st->alpha;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s[i] = fma(-alpha, t0[i], r[i]);
}

// x += alpha p + omega s; r = s - omega t1; ||r||^2, <r,r0*>; then beta / restart bookkeeping   (bicgstab.hpp:458-474)
static __global__ void __launch_bounds__(VEC_THREADS)
pbicg_xr_kernel(long long n, real *x, const real *p, const real *s, real *r, const real *t1, const real *r0,
                SolverState *st, real *partials, unsigned int *ticket)
{
  __shared__ real s_red[64];
  if (st->done != VCL_RUNNING || st->need_restart) return;
  const real alpha = st->alpha, omega = st->omega;
  real acc[2] = {0.0, 0.0};
  const long long npairs = aligned16(x, p, s, r, t1, r0) ? (n >> 1) : 0;
  for (long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x; pi < npairs; pi += (long long)gridDim.x * blockDim.x)
  {
    const long long k = 2 * pi;
    const real2 vs = ld2(s, k), vp = ld2(p, k), vt = ld2(t1, k), v0 = ld2(r0, k);
    real2 vx = ld2(x, k), vr;
    vx.x += alpha * vp.x + omega * vs.x;   vx.y += alpha * vp.y + omega * vs.y;
    vr.x = fma(-omega, vt.x, vs.x);        vr.y = fma(-omega, vt.y, vs.y);
    acc[0] = fma(vr.x, vr.x, acc[0]);      acc[0] = fma(vr.y, vr.y, acc[0]);
    acc[1] = fma(vr.x, v0.x, acc[1]);      acc[1] = fma(vr.y, v0.y, acc[1]);
    st2(x, k, vx); st2(r, k, vr);
  }
  for (long long i = 2 * npairs + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
  {
    const real vs = s[i];
    x[i] += alpha * p[i] + omega * vs;
    const real vr = fma(-omega, t1[i], vs);
    r[i] = vr;
    acc[0] = fma(vr, vr, acc[0]);
    acc[1] = fma(vr, r0[i], acc[1]);
  }
  if (grid_sum_last_block<2>(acc, partials, ticket, s_red) && threadIdx.x == 0)
  {
    const int i = st->iters;            // index of the iteration just finished
    st->iters = i + 1;
    const real res = sqrt(acc[0]);
    st->residual_norm = res;
    st->est = fabs(res / st->norm_rhs);
    if (res / st->norm_rhs < st->tol || res < st->abs_tol) { st->done = VCL_CONVERGED; return; }
    const real new_ip = acc[1];
    st->beta = new_ip / st->ip_rr0 * alpha / omega;
    st->ip_rr0 = new_ip;
    if (new_ip == 0.0 || omega == 0.0 || i - st->last_restart > st->restart_every) st->need_restart = 1;
    if (st->iters >= st->maxit) st->done = VCL_MAXIT;
  }
}

// p -= omega t0; p = r + beta p                               (bicgstab.hpp:479-480)
static __global__ void __launch_bounds__(VEC_THREADS)
pbicg_p_kernel(long long n, real *p, const real *r, const real *t0, const SolverState *st)
{
  if (st->done != VCL_RUNNING || st->need_restart) return;
  const real beta = st->beta, omega = st->omega;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = fma(beta, fma(-omega, t0[i], p[i]), r[i]);
}

// restart: r = (b - r) [/ diag]; p = r; r0 = r; ip_rr0 = ||r||^2     (bicgstab.hpp:430-442; r holds A*x on entry)
template<bool JACOBI>
static __global__ void __launch_bounds__(VEC_THREADS)
pbicg_restart_kernel(long long n, const real *b, real *r, real *p, real *r0, const real *diag,
                     SolverState *st, real *partials, unsigned int *ticket)
{
  __shared__ real s_red[32];
  real acc[1] = {0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
  {
    real v = b[i] - r[i];
    if (JACOBI) v = v / diag[i];
    r[i] = v; p[i] = v; r0[i] = v;
    acc[0] = fma(v, v, acc[0]);
  }
  if (grid_sum_last_block<1>(acc, partials, ticket, s_red) && threadIdx.x == 0)
  {
    const real nrm = sqrt(acc[0]);
    st->ip_rr0 = nrm * nrm;
    st->need_restart = 0;
    st->last_restart = st->iters;
  }
}

// ------------------------------------------------------------------------------------------------
// GMRES (classical Gram-Schmidt, gmres.hpp:241-284)
// ------------------------------------------------------------------------------------------------
// stage 1: h_j = <v_j, v_k>, j < k, ALL k vectors in one pass over the basis (the reference sweeps 7 vectors at a time and
// re-reads v_k for every sweep, cuda/iterative_operations.hpp:1690-1735).
// Work split: a 2-D grid.  blockIdx.y selects a group of <= GS1_COLS columns, blockIdx.x a persistent share of the rows.
// Every thread of a CTA does the same work (its 16-byte slice of v_k against the <= 8 columns of the group, 8 accumulators),
// so the warps of a CTA are balanced for every k -- a first version dealt columns to warps and lost up to 44 % at
// k = 9, 17, 25 (profiles/ncu_summary_r1b.md: gs1<4> at 3.5 TB/s).  The column groups of one row share are resident at
// the same time and re-read the same slice of v_k, which L2 serves: DRAM traffic stays (k+1)*8*n bytes.
#define GS1_COLS 8
static __global__ void __launch_bounds__(VEC_THREADS)
gmres_gs1_kernel(const real *basis, long long n, long long isz, int k, real *out_h, int out_stride,
                 real *partials, unsigned int *ticket)
{
  __shared__ real s_part[8][GS1_COLS];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c0 = blockIdx.y * GS1_COLS;
  const int nc = min(GS1_COLS, k - c0);
  const real *vk = basis + (size_t)k * isz;
  const real *col = basis + (size_t)c0 * isz;
  real acc[GS1_COLS];
#pragma unroll
  for (int q = 0; q < GS1_COLS; ++q) acc[q] = 0.0;
  const long long npairs = n >> 1;
  for (long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x; pi < npairs; pi += (long long)gridDim.x * blockDim.x)
  {
    const real2 v = ld2(vk, 2 * pi);
    real2 a[GS1_COLS];
#pragma unroll
    for (int q = 0; q < GS1_COLS; ++q)
      if (q < nc) a[q] = ld2(col + (size_t)q * isz, 2 * pi);
#pragma unroll
    for (int q = 0; q < GS1_COLS; ++q)
      if (q < nc) { acc[q] = fma(a[q].x, v.x, acc[q]); acc[q] = fma(a[q].y, v.y, acc[q]); }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
  {
    const real v = vk[n - 1];
#pragma unroll
    for (int q = 0; q < GS1_COLS; ++q)
      if (q < nc) acc[q] = fma(col[(size_t)q * isz + n - 1], v, acc[q]);
  }
#pragma unroll
  for (int q = 0; q < GS1_COLS; ++q)
  {
    const real t = warp_sum(acc[q]);
    if (lane == 0) s_part[w][q] = t;
  }
  __syncthreads();
  if ((int)threadIdx.x < nc)
  {
    real t = 0.0;
    for (int r = 0; r < 8; ++r) t += s_part[r][threadIdx.x];
    partials[(size_t)(c0 + threadIdx.x) * VCL_MAX_BLOCKS + blockIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA: warp w finishes columns w, w+8, ... in a fixed order
  for (int j = w; j < k; j += 8)
  {
    real t = 0.0;
    for (unsigned int i = lane; i < gridDim.x; i += 32) t += __ldcg(partials + (size_t)j * VCL_MAX_BLOCKS + i);
    t = warp_sum(t);
    if (lane == 0) out_h[(size_t)j * out_stride] = t;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// stage 2: v_k -= sum_j h_j v_j; R[j + k*m] = h_j; ||v_k||^2       (host_based/iterative_operations.hpp:852-893)
static __global__ void __launch_bounds__(VEC_THREADS)
gmres_gs2_kernel(real *basis, long long n, long long isz, int k, const real *h, int h_stride,
                 real *R, int krylov_dim, real *out_norm_sq, real *partials, unsigned int *ticket)
{
  __shared__ real s_red[32];
  __shared__ real s_h[VCL_GMRES_MAX_KRYLOV];
  for (int j = threadIdx.x; j < k; j += blockDim.x) s_h[j] = h[(size_t)j * h_stride];
  __syncthreads();
  real *vk = basis + (size_t)k * isz;
  real acc[1] = {0.0};
  const long long npairs = n >> 1;
  for (long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x; pi < npairs; pi += (long long)gridDim.x * blockDim.x)
  {
    real2 v = ld2(vk, 2 * pi);
    int j = 0;
    for (; j + 4 <= k; j += 4)
    {
      const real2 a0 = ld2(basis + (size_t)j * isz, 2 * pi), a1 = ld2(basis + (size_t)(j + 1) * isz, 2 * pi);
      const real2 a2 = ld2(basis + (size_t)(j + 2) * isz, 2 * pi), a3 = ld2(basis + (size_t)(j + 3) * isz, 2 * pi);
      v.x = fma(-s_h[j], a0.x, v.x);     v.y = fma(-s_h[j], a0.y, v.y);
      v.x = fma(-s_h[j + 1], a1.x, v.x); v.y = fma(-s_h[j + 1], a1.y, v.y);
      v.x = fma(-s_h[j + 2], a2.x, v.x); v.y = fma(-s_h[j + 2], a2.y, v.y);
      v.x = fma(-s_h[j + 3], a3.x, v.x); v.y = fma(-s_h[j + 3], a3.y, v.y);
    }
    for (; j < k; ++j)
    {
      const real2 a0 = ld2(basis + (size_t)j * isz, 2 * pi);
      v.x = fma(-s_h[j], a0.x, v.x); v.y = fma(-s_h[j], a0.y, v.y);
    }
    acc[0] = fma(v.x, v.x, acc[0]); acc[0] = fma(v.y, v.y, acc[0]);
    st2(vk, 2 * pi, v);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
  {
    real v = vk[n - 1];
    for (int j = 0; j < k; ++j) v = fma(-s_h[j], basis[(size_t)j * isz + n - 1], v);
    acc[0] = fma(v, v, acc[0]);
    vk[n - 1] = v;
  }
  if (grid_sum_last_block<1>(acc, partials, ticket, s_red))
  {
    for (int j = threadIdx.x; j < k; j += blockDim.x) R[(size_t)j + (size_t)k * krylov_dim] = s_h[j];
    if (threadIdx.x == 0) *out_norm_sq = acc[0];
  }
}

// normalize: R[off] = ||v_k||; v_k /= ||v_k||; xi_k = <r, v_k>      (host_based/iterative_operations.hpp:733-778)
static __global__ void __launch_bounds__(VEC_THREADS)
gmres_normalize_kernel(long long n, real *vk, const real *res, real *R, int offset_in_R, const real *in_norm_sq,
                       real *out_r_dot_vk, real *partials, unsigned int *ticket)
{
  __shared__ real s_red[32];
  const real nrm = sqrt(*in_norm_sq);
  if (blockIdx.x == 0 && threadIdx.x == 0) R[offset_in_R] = nrm;
  real acc[1] = {0.0};
  const bool vec = ((reinterpret_cast<uintptr_t>(vk) | reinterpret_cast<uintptr_t>(res)) & 15u) == 0u;
  const long long npairs = vec ? (n >> 1) : 0;
  for (long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x; pi < npairs; pi += (long long)gridDim.x * blockDim.x)
  {
    real2 v = ld2(vk, 2 * pi); const real2 rr = ld2(res, 2 * pi);
    v.x = v.x / nrm; v.y = v.y / nrm;
    acc[0] = fma(rr.x, v.x, acc[0]); acc[0] = fma(rr.y, v.y, acc[0]);
    st2(vk, 2 * pi, v);
  }
  for (long long i = 2 * npairs + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
  {
    const real v = vk[i] / nrm;
    acc[0] = fma(res[i], v, acc[0]);
    vk[i] = v;
  }
  if (grid_sum_last_block<1>(acc, partials, ticket, s_red) && threadIdx.x == 0) *out_r_dot_vk = acc[0];
}

// x += c_0 r + sum_{j=1}^{k-1} c_j v_{j-1}                         (host_based/iterative_operations.hpp:895-922)
static __global__ void __launch_bounds__(VEC_THREADS)
gmres_update_kernel(long long n, real *x, const real *res, const real *basis, long long isz, const real *coef, int k)
{
  __shared__ real s_c[VCL_GMRES_MAX_KRYLOV];
  for (int j = threadIdx.x; j < max(k, 1); j += blockDim.x) s_c[j] = coef[j];
  __syncthreads();
  const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(basis)) & 15u) == 0u && (isz & 1) == 0;
  const long long npairs = vec ? (n >> 1) : 0;
  for (long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x; pi < npairs; pi += (long long)gridDim.x * blockDim.x)
  {
    real2 v = ld2(x, 2 * pi); const real2 rr = ld2(res, 2 * pi);
    v.x = fma(s_c[0], rr.x, v.x); v.y = fma(s_c[0], rr.y, v.y);
    for (int j = 1; j < k; ++j)
    {
      const real2 a = ld2(basis + (size_t)(j - 1) * isz, 2 * pi);
      v.x = fma(s_c[j], a.x, v.x); v.y = fma(s_c[j], a.y, v.y);
    }
    st2(x, 2 * pi, v);
  }
  for (long long i = 2 * npairs + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
  {
    real v = x[i];
    v = fma(s_c[0], res[i], v);
    for (int j = 1; j < k; ++j) v = fma(s_c[j], basis[(size_t)(j - 1) * isz + i], v);
    x[i] = v;
  }
}
} // namespace VCL_NS
