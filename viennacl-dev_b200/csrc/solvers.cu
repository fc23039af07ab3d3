// solvers.cu -- solver drivers (CG / BiCGStab / GMRES) and the per-step C-ABI entry points.
// Drivers restate linalg/cg.hpp:128-187, bicgstab.hpp:97-215 and :398-489, gmres.hpp:181-367 with the iteration loop kept
// next to the kernels: scalars live in device memory (SolverState), iterations are enqueued in batches, and the host looks
// at the state once per batch (with a monitor callback installed: once per iteration, as the reference's contract demands).
#include "fused_kernels.cuh"
#include "launch.cuh"
#include "nvtx.cuh"
#include "gmres_host.cuh"
#include "gmres_launch.cuh"
#include "persistent.cuh"
#include "blas1.cuh"
#include <cmath>
#include <algorithm>

namespace VCL_NS
{
namespace {

struct MatOp
{
  int fmt;                 // 0 CSR, 1 SELL, 2 ELL / HYB (plain ELL = HYB without tail)
  ViennaCLCUDADcsr csr;
  ViennaCLCUDADsell sell;
  ViennaCLCUDADhyb hyb;
  int rows() const { return fmt == 0 ? csr.rows : (fmt == 1 ? sell.rows : hyb.ell.rows); }
  int cols() const { return fmt == 0 ? csr.cols : (fmt == 1 ? sell.cols : hyb.ell.cols); }
};

template<class Epi>
ViennaCLStatus launch_prod(ViennaCLBackend b, const MatOp &A, const real *x, Epi epi)
{
  XVec xv = make_xvec(x, 0, 1);
  if (A.fmt == 0) return vcl_launch_csr(b, A.csr, xv, epi);
  if (A.fmt == 1) return vcl_launch_sell(b, A.sell, xv, epi);
  return vcl_launch_ell(b, A.hyb, xv, epi);
}

ViennaCLStatus plain_prod(ViennaCLBackend b, const MatOp &A, const real *x, real *y)
{
  EpiAxpby epi = {y, 0, 1, 1.0, 0.0};
  return launch_prod(b, A, x, epi);
}

int vec_grid(ViennaCLBackend b, long long n)
{
  long long want = (n / 2 + VEC_THREADS - 1) / VEC_THREADS;
  return (int)std::max(1LL, std::min(want, (long long)std::min(b->sm_count * 8, VCL_MAX_BLOCKS)));
}

ViennaCLStatus check_matrix(ViennaCLBackend b, const MatOp &A)
{
  VCL_REQUIRE(b, A.rows() >= 0 && A.rows() == A.cols(), "solvers need a square matrix");
  if (A.fmt == 0) VCL_REQUIRE(b, A.rows() == 0 || (A.csr.row_ptr && (A.csr.nnz == 0 || (A.csr.col_idx && A.csr.values))), "null CSR array");
  else if (A.fmt == 1) VCL_REQUIRE(b, A.rows() == 0 || (A.sell.columns_per_block && A.sell.block_start && A.sell.rows_per_block > 0), "bad SELL matrix");
  else VCL_REQUIRE(b, A.rows() == 0 || (A.hyb.ell.maxnnz >= 0 && A.hyb.ell.internal_rows >= A.hyb.ell.rows &&
                                        (A.hyb.ell.maxnnz == 0 || (A.hyb.ell.coords && A.hyb.ell.elements)) &&
                                        (!A.hyb.csr_rows || (A.hyb.csr_cols && A.hyb.csr_elements))), "bad ELL / HYB matrix");
  return ViennaCLSuccess;
}

ViennaCLStatus check_matrix_any(ViennaCLBackend b, const MatOp &A)       // like check_matrix, rectangular allowed (plain products)
{
  VCL_REQUIRE(b, A.rows() >= 0 && A.cols() >= 0, "negative size");
  VCL_REQUIRE(b, A.rows() == 0 || (A.hyb.ell.maxnnz >= 0 && A.hyb.ell.internal_rows >= A.hyb.ell.rows &&
                                   (A.hyb.ell.maxnnz == 0 || (A.hyb.ell.coords && A.hyb.ell.elements)) &&
                                   (!A.hyb.csr_rows || (A.hyb.csr_cols && A.hyb.csr_elements))), "bad ELL / HYB matrix");
  return ViennaCLSuccess;
}

// carve 256-byte aligned sub-buffers out of the backend workspace
struct Carver
{
  char *base; size_t off;
  explicit Carver(void *p) : base((char*)p), off(0) {}
  real *take(size_t n) { real *p = (real*)(base + off); off += ((n * sizeof(real) + 255) / 256) * 256; return p; }
  static size_t need(size_t n) { return ((n * sizeof(real) + 255) / 256) * 256; }
};

ViennaCLStatus push_state(ViennaCLBackend b)
{
  VCL_CUDA(b, cudaMemcpyAsync(VCL_DSTATE(b), VCL_HSTATE(b), sizeof(SolverState), cudaMemcpyHostToDevice, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));   // hstate is reused as the read-back mirror
  return ViennaCLSuccess;
}

ViennaCLStatus pull_state(ViennaCLBackend b)
{
  VCL_CUDA(b, cudaMemcpyAsync(VCL_HSTATE(b), VCL_DSTATE(b), sizeof(SolverState), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  return ViennaCLSuccess;
}

// diagonal preconditioners: which row_info option yields the scaling vector (cuda/sparse_matrix_operations.hpp:53-119:
// 0 inf-norm, 1 1-norm, 2 2-norm, 3 diagonal); -1: not a diagonal preconditioner
static inline int diag_precond_option(int precond)
{
  switch (precond)
  {
  case ViennaCLB200PrecondJacobi: return 3;
  case ViennaCLB200PrecondRowScalingInf: return 0;
  case ViennaCLB200PrecondRowScaling1: return 1;
  case ViennaCLB200PrecondRowScaling2: return 2;
  default: return -1;
  }
}

const int kBatch = 32;   // iterations enqueued between two looks at the device state

// Use the persistent cooperative kernel?  CSR with a row-block plan and 16-byte aligned arrays (the TMA path), a device that
// supports cooperative launches, and a system small enough that fixed latencies matter (option "persistent_rows" of the
// handle, default 10M rows -- measured: ahead up to 200^3, level at 256^3; 0 disables).
static bool persistent_cg_wanted(ViennaCLBackend b, const ViennaCLCUDADcsr &A, long long n, int row_limit_divisor = 1)
{
  const long long max_rows = b->persistent_rows >= 0 ? b->persistent_rows : 10000000LL;
  return b->coop_launch == 1 && n <= max_rows / row_limit_divisor && A.row_blocks && A.num_blocks > 0 && vcl_aligned16(A.values) && vcl_aligned16(A.col_idx) &&
         vcl_plan_ok(b, A.row_ptr, A.rows, A.nnz, A.row_blocks, A.num_blocks);
}

// ------------------------------------------------------------------------------------------------
// CG with Jacobi preconditioner: the reference runs its generic PCG (cg.hpp:257-322: SpMV + element_div + ~6 BLAS-1
// launches + 2 blocking reductions per iteration).  Here: single-reduction (Chronopoulos/Gear) PCG, 2 kernels per
// iteration -- pcg_update_kernel (all vector updates, u = r ./ diag, <r,u>) and the fused SpMV w = A u with <w,u>,
// whose last CTA advances alpha/beta/convergence on the device.  12*nnz + 116*N bytes per iteration.
// ------------------------------------------------------------------------------------------------
ViennaCLStatus pcg_jacobi(ViennaCLBackend b, const MatOp &A, const real *rhs, real *x, ViennaCLB200SolverTag *tag)
{
  VCL_RANGE("vcl:pcg_jacobi");
  VCL_REQUIRE(b, A.fmt == 0, "Jacobi needs the CSR matrix (row_info, linalg/sparse_matrix_operations.hpp:48-74)");
  const long long n = A.rows();
  VCL_TRY(vcl_ws_reserve(b, 6 * Carver::need(n)));
  Carver cv(b->ws);
  real *r = cv.take(n), *u = cv.take(n), *w = cv.take(n), *p = cv.take(n), *s = cv.take(n), *diag = cv.take(n);
  const int grid = vec_grid(b, n);

  VCL_TRY(ViennaCLCUDADcsr_row_info(b, (int)n, A.csr.row_ptr, A.csr.col_idx, A.csr.values, diag, diag_precond_option(tag->precond)));
  VCL_CUDA(b, cudaMemsetAsync(x, 0, sizeof(real) * n, b->stream));
  VCL_CUDA(b, cudaMemsetAsync(p, 0, sizeof(real) * n, b->stream));
  VCL_CUDA(b, cudaMemsetAsync(s, 0, sizeof(real) * n, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(r, rhs, sizeof(real) * n, cudaMemcpyDeviceToDevice, b->stream));
  pcg_init_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, r, u, diag, VCL_PARTIALS(b), b->tickets, VCL_DSCAL(b) + 0);
  VCL_LAUNCHED(b, "pcg_init_kernel");
  VCL_TRY(plain_prod(b, A, u, w));
  VCL_TRY(vcl_dot_async(b, n, w, 0, 1, u, 0, 1, VCL_DSCAL(b) + 1));
  VCL_CUDA(b, cudaMemcpyAsync(VCL_HSCAL(b), VCL_DSCAL(b), 2 * sizeof(real), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  const real gamma0 = VCL_HSCAL(b)[0], delta0 = VCL_HSCAL(b)[1];
  if (std::fabs(gamma0) <= tag->abs_tolerance * tag->abs_tolerance) return ViennaCLSuccess;       // cg.hpp:286-287

  SolverState *h = VCL_HSTATE(b);
  std::memset(h, 0, sizeof(SolverState));
  h->alpha = gamma0 / delta0; h->beta = 0.0; h->ip_rr0 = gamma0; h->norm_rhs_sq = gamma0; h->norm_rhs = std::sqrt(std::fabs(gamma0));
  h->tol = tag->tolerance; h->abs_tol = tag->abs_tolerance; h->maxit = tag->max_iterations; h->sums[0] = gamma0;
  VCL_TRY(push_state(b));
  SolverState *st = VCL_DSTATE(b);

  const int batch = tag->monitor ? 1 : kBatch;
  int coop_grid = 0;                                       // persistent form for small / medium systems, see cg_solve
  if (persistent_cg_wanted(b, A.csr, n, 3))                // 7 streamed vectors per update: level with two kernels from ~2M rows (128^3), behind at 256^3
  {
    const int occ = vcl_occupancy(b, pcg_persistent_kernel, CSR_BLOCK_THREADS, CSR_SMEM_BYTES);
    coop_grid = std::max(1, std::min(b->sm_count * occ, std::max(A.csr.num_blocks, vcl_div_up(n, 2 * CSR_BLOCK_THREADS))));
  }
  int launched = 0;
  while (launched < tag->max_iterations)
  {
    const int nb = std::min(batch, tag->max_iterations - launched);
    VCL_RANGE("vcl:batch");
    if (coop_grid > 0)
    {
      CsrDev d = {A.csr.rows, (u32)A.csr.nnz, A.csr.row_ptr, A.csr.col_idx, A.csr.values, A.csr.row_blocks, A.csr.row_blocks + 1, A.csr.num_blocks};
      d.l2_mode = vcl_l2_mode(b, A.csr.nnz, A.csr.rows);
      XVec xv = make_xvec(u, 0, 1);
      long long nn = n; int iters_arg = nb;
      real *partials = VCL_PARTIALS(b);
      const real *cdiag = diag;
      void *args[] = {&d, &xv, &nn, &x, &r, &u, &w, &p, &s, &cdiag, &st, &partials, &iters_arg};
      const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)pcg_persistent_kernel, dim3(coop_grid), dim3(CSR_BLOCK_THREADS), args,
                                                         (size_t)CSR_SMEM_BYTES, b->stream);
      if (ce == cudaErrorCooperativeLaunchTooLarge || ce == cudaErrorLaunchOutOfResources) { (void)cudaGetLastError(); coop_grid = 0; }
      else { VCL_CUDA(b, ce); VCL_LAUNCHED(b, "pcg_persistent_kernel"); }
    }
    if (coop_grid == 0)
    for (int k = 0; k < nb; ++k)
    {
      pcg_update_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, x, r, u, w, p, s, diag, st, VCL_PARTIALS(b), b->tickets, &st->sums[0]);
      VCL_LAUNCHED(b, "pcg_update_kernel");
      EpiFused<STEP_PCG, false, false> epi = {w, u, nullptr, nullptr, VCL_PARTIALS(b), b->tickets, st, nullptr, nullptr, nullptr, {0.0, 0.0, 0.0}, nullptr};
      VCL_TRY(launch_prod(b, A, u, epi));
    }
    launched += nb;
    VCL_TRY(pull_state(b));
    if (tag->monitor && tag->monitor(x, h->est, tag->monitor_user)) break;
    if (h->done != VCL_RUNNING) break;
  }
  tag->iters = h->iters;
  tag->error = std::sqrt(std::fabs(h->sums[0] / gamma0));                            // cg.hpp:317
  return ViennaCLSuccess;
}

// ------------------------------------------------------------------------------------------------
// CG  (cg.hpp:128-187)
// ------------------------------------------------------------------------------------------------
ViennaCLStatus cg_solve(ViennaCLBackend b, const MatOp &A, const real *rhs, real *x, ViennaCLB200SolverTag *tag)
{
  VCL_RANGE("vcl:cg");
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, tag != nullptr, "null tag");
  VCL_TRY(check_matrix(b, A));
  const long long n = A.rows();
  tag->iters = 0; tag->error = 0.0;
  if (n == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, rhs && x, "null vector");
  VCL_CUDA(b, cudaSetDevice(b->device));
  if (diag_precond_option(tag->precond) >= 0) return pcg_jacobi(b, A, rhs, x, tag);
  VCL_REQUIRE(b, tag->precond == ViennaCLB200PrecondNone, "CG: unknown preconditioner id");
  // Small and medium CSR systems: whole iterations inside one cooperative kernel (persistent.cuh).  The one-pass form keeps a
  // second set of r / p / Ap (double buffering), hence six work vectors instead of three.
  const bool want_coop = A.fmt == 0 && persistent_cg_wanted(b, A.csr, n);
  VCL_TRY(vcl_ws_reserve(b, (want_coop ? 6 : 3) * Carver::need(n)));
  Carver cv(b->ws);
  real *r = cv.take(n), *p = cv.take(n), *Ap = cv.take(n);
  real *r2 = want_coop ? cv.take(n) : nullptr, *p2 = want_coop ? cv.take(n) : nullptr, *Ap2 = want_coop ? cv.take(n) : nullptr;

  VCL_CUDA(b, cudaMemsetAsync(x, 0, sizeof(real) * n, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(r, rhs, sizeof(real) * n, cudaMemcpyDeviceToDevice, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(p, rhs, sizeof(real) * n, cudaMemcpyDeviceToDevice, b->stream));
  VCL_TRY(plain_prod(b, A, p, Ap));
  VCL_TRY(vcl_dot_async(b, n, r, 0, 1, r, 0, 1, VCL_DSCAL(b) + 0));
  VCL_TRY(vcl_dot_async(b, n, p, 0, 1, Ap, 0, 1, VCL_DSCAL(b) + 1));
  VCL_TRY(vcl_dot_async(b, n, Ap, 0, 1, Ap, 0, 1, VCL_DSCAL(b) + 2));
  VCL_CUDA(b, cudaMemcpyAsync(VCL_HSCAL(b), VCL_DSCAL(b), 3 * sizeof(real), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));

  real norm_rhs_squared = std::sqrt(VCL_HSCAL(b)[0]); norm_rhs_squared *= norm_rhs_squared;          // cg.hpp:147
  if (norm_rhs_squared <= tag->abs_tolerance * tag->abs_tolerance) return ViennaCLSuccess;         // cg.hpp:149-150
  const real rr = norm_rhs_squared;
  const real alpha = rr / VCL_HSCAL(b)[1];
  real beta = std::sqrt(VCL_HSCAL(b)[2]); beta = (alpha * alpha * beta * beta - rr) / rr;            // cg.hpp:153-154

  SolverState *h = VCL_HSTATE(b);
  std::memset(h, 0, sizeof(SolverState));
  h->alpha = alpha; h->beta = beta; h->norm_rhs_sq = norm_rhs_squared; h->norm_rhs = std::sqrt(norm_rhs_squared);
  h->tol = tag->tolerance; h->abs_tol = tag->abs_tolerance; h->maxit = tag->max_iterations; h->sums[0] = rr;
  VCL_TRY(push_state(b));
  SolverState *st = VCL_DSTATE(b);

  const int grid = vec_grid(b, n);
  const int batch = tag->monitor ? 1 : kBatch;
  // Persistent forms (persistent.cuh): "onepass" (default) = one phase and one grid barrier per iteration, the product recomputes
  // the updated p on the fly; "twophase" = update phase, barrier, product phase, barrier (option "persistent_cg_form" = 2).
  // Large systems are bound by HBM and keep the two-kernel form.
  int coop_grid = 0;
  // form 0 (default) = by size: the one-pass form wins while fixed latencies dominate (512^2: 10.5 against 13.5 us per iteration),
  // is level at 1M rows and loses once the three gathers per entry cost more L2 traffic than the saved barrier (2048^2: 117
  // against 104 us; profiles/cg_forms_r2b.log)
  const bool onepass = b->persistent_cg_form == 0 ? n <= 600000 : b->persistent_cg_form != 2;
  const void *onepass_fn = b->persistent_cg_form == 3 ? (const void*)cg_onepass_kernel<3> : (const void*)cg_onepass_kernel<2>;
  if (want_coop)
  {
    const int occ = !onepass ? vcl_occupancy(b, cg_persistent_kernel, CSR_BLOCK_THREADS, CSR_SMEM_BYTES)
                             : (b->persistent_cg_form == 3 ? vcl_occupancy(b, cg_onepass_kernel<3>, CSR_BLOCK_THREADS, CSR_SMEM_BYTES)
                                                           : vcl_occupancy(b, cg_onepass_kernel<2>, CSR_BLOCK_THREADS, CSR_SMEM_BYTES));
    coop_grid = std::max(1, std::min(b->sm_count * occ, std::max(A.csr.num_blocks, vcl_div_up(n, 2 * CSR_BLOCK_THREADS))));
    if (onepass) coop_grid = std::max(1, std::min(b->sm_count * occ, A.csr.num_blocks));
  }
  int launched = 0;
  while (launched < tag->max_iterations)
  {
    const int nb = std::min(batch, tag->max_iterations - launched);
    VCL_RANGE("vcl:batch");
    if (coop_grid > 0)
    {
      CsrDev d = {A.csr.rows, (u32)A.csr.nnz, A.csr.row_ptr, A.csr.col_idx, A.csr.values, A.csr.row_blocks, A.csr.row_blocks + 1, A.csr.num_blocks};
      d.l2_mode = vcl_l2_mode(b, A.csr.nnz, A.csr.rows);
      XVec xv = make_xvec(p, 0, 1);
      long long nn = n; int iters_arg = nb;
      real *partials = VCL_PARTIALS(b);
      void *args2[] = {&d, &xv, &nn, &x, &p, &r, &Ap, &st, &partials, &iters_arg};
      void *args1[] = {&d, &x, &r, &p, &Ap, &r2, &p2, &Ap2, &st, &partials, &iters_arg};
      const cudaError_t ce = onepass ? cudaLaunchCooperativeKernel(onepass_fn, dim3(coop_grid), dim3(CSR_BLOCK_THREADS), args1,
                                                                   (size_t)CSR_SMEM_BYTES, b->stream)
                                     : cudaLaunchCooperativeKernel((const void*)cg_persistent_kernel, dim3(coop_grid), dim3(CSR_BLOCK_THREADS), args2,
                                                                   (size_t)CSR_SMEM_BYTES, b->stream);
      if (ce == cudaErrorCooperativeLaunchTooLarge || ce == cudaErrorLaunchOutOfResources)
      {
        (void)cudaGetLastError();                          // the SMs are shared with another client (MPS, a second stream): two-kernel form
        coop_grid = 0;
        if (h->pad0) { std::swap(r, r2); std::swap(p, p2); std::swap(Ap, Ap2); }      // the current r / p / Ap are in the second set
      }
      else
      {
        VCL_CUDA(b, ce);
        VCL_LAUNCHED(b, onepass ? "cg_onepass_kernel" : "cg_persistent_kernel");
      }
    }
    if (coop_grid == 0)
    for (int k = 0; k < nb; ++k)
    {
      cg_update_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, x, p, r, Ap, 0.0, 0.0, st, VCL_PARTIALS(b), b->tickets, &st->sums[0]);
      VCL_LAUNCHED(b, "cg_update_kernel");
      EpiFused<STEP_CG, false, false> epi = {Ap, p, nullptr, nullptr, VCL_PARTIALS(b), b->tickets, st, &st->sums[1], &st->sums[2], nullptr, {0.0, 0.0, 0.0}, nullptr};
      VCL_TRY(launch_prod(b, A, p, epi));
    }
    launched += nb;
    VCL_TRY(pull_state(b));
    if (tag->monitor && tag->monitor(x, h->est, tag->monitor_user)) break;          // cg.hpp:174 (monitor first, then the test)
    if (h->done != VCL_RUNNING) break;
  }
  tag->iters = h->iters;
  tag->error = std::sqrt(std::fabs(h->sums[0]) / norm_rhs_squared);                 // cg.hpp:184
  return ViennaCLSuccess;
}

// ------------------------------------------------------------------------------------------------
// Pipelined BiCGStab  (bicgstab.hpp:97-215)
// ------------------------------------------------------------------------------------------------
ViennaCLStatus bicgstab_pipelined(ViennaCLBackend b, const MatOp &A, const real *rhs, real *x, ViennaCLB200SolverTag *tag)
{
  VCL_RANGE("vcl:bicgstab_pipelined");
  const long long n = A.rows();
  VCL_TRY(vcl_ws_reserve(b, 6 * Carver::need(n)));
  Carver cv(b->ws);
  real *r = cv.take(n), *p = cv.take(n), *r0 = cv.take(n), *Ap = cv.take(n), *s = cv.take(n), *As = cv.take(n);

  VCL_CUDA(b, cudaMemsetAsync(x, 0, sizeof(real) * n, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(r, rhs, sizeof(real) * n, cudaMemcpyDeviceToDevice, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(p, rhs, sizeof(real) * n, cudaMemcpyDeviceToDevice, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(r0, rhs, sizeof(real) * n, cudaMemcpyDeviceToDevice, b->stream));
  real ss = 0.0;
  VCL_TRY(vcl_dot_host(b, n, r, 0, 1, r, 0, 1, &ss));
  const real norm_rhs = std::sqrt(ss);
  if (norm_rhs <= tag->abs_tolerance) return ViennaCLSuccess;                       // bicgstab.hpp:140-141

  SolverState *h = VCL_HSTATE(b);
  std::memset(h, 0, sizeof(SolverState));
  h->norm_rhs = norm_rhs; h->norm_rhs_sq = norm_rhs * norm_rhs; h->residual_norm = norm_rhs;
  h->tol = tag->tolerance; h->abs_tol = tag->abs_tolerance; h->maxit = tag->max_iterations;
  h->sums[0] = norm_rhs * norm_rhs;                                                 // bicgstab.hpp:131
  VCL_TRY(push_state(b));
  SolverState *st = VCL_DSTATE(b);

  const int grid = vec_grid(b, n);
  const int batch = tag->monitor ? 1 : kBatch;
  int launched = 0;
  bool stopped = false;
  int coop_grid = 0;                                       // persistent form for small / medium CSR systems, see cg_solve
  if (A.fmt == 0 && !tag->monitor && persistent_cg_wanted(b, A.csr, n))
  {
    const int occ = vcl_occupancy(b, bicgstab_persistent_kernel, CSR_BLOCK_THREADS, CSR_SMEM_BYTES);
    coop_grid = std::max(1, std::min(b->sm_count * occ, std::max(A.csr.num_blocks, vcl_div_up(n, 2 * CSR_BLOCK_THREADS))));
  }
  while (launched < tag->max_iterations && !stopped)
  {
    const int nb = std::min(batch, tag->max_iterations - launched);
    VCL_RANGE("vcl:batch");
    if (coop_grid > 0)
    {
      CsrDev d = {A.csr.rows, (u32)A.csr.nnz, A.csr.row_ptr, A.csr.col_idx, A.csr.values, A.csr.row_blocks, A.csr.row_blocks + 1, A.csr.num_blocks};
      d.l2_mode = vcl_l2_mode(b, A.csr.nnz, A.csr.rows);
      long long nn = n; int iters_arg = nb;
      real *partials = VCL_PARTIALS(b);
      const real *cr0 = r0;
      void *args[] = {&d, &nn, &x, &r, &p, &cr0, &Ap, &s, &As, &st, &partials, &iters_arg};
      const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)bicgstab_persistent_kernel, dim3(coop_grid), dim3(CSR_BLOCK_THREADS), args,
                                                         (size_t)CSR_SMEM_BYTES, b->stream);
      if (ce == cudaErrorCooperativeLaunchTooLarge || ce == cudaErrorLaunchOutOfResources) { (void)cudaGetLastError(); coop_grid = 0; }
      else { VCL_CUDA(b, ce); VCL_LAUNCHED(b, "bicgstab_persistent_kernel"); }
    }
    if (coop_grid == 0)
    for (int k = 0; k < nb; ++k)
    {
      EpiFused<STEP_NONE, true, false> e1 = {Ap, p, r0, nullptr, VCL_PARTIALS(b), b->tickets, st, &st->sums[1], &st->sums[2], &st->sums[3], {0.0, 0.0, 0.0}, nullptr};
      VCL_TRY(launch_prod(b, A, p, e1));
      bicgstab_update_s_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, s, r, Ap, &st->sums[0], &st->sums[3], st, VCL_PARTIALS(b), b->tickets, &st->sums[5]);
      VCL_LAUNCHED(b, "bicgstab_update_s_kernel");
      EpiFused<STEP_BICGSTAB, true, false> e2 = {As, s, r0, nullptr, VCL_PARTIALS(b), b->tickets, st, &st->sums[1], &st->sums[2], &st->sums[4], {0.0, 0.0, 0.0}, nullptr};
      VCL_TRY(launch_prod(b, A, s, e2));
      if (tag->monitor)
      {
        // the reference shows the monitor the iterate BEFORE this step's vector update (bicgstab.hpp:196-206)
        VCL_TRY(pull_state(b));
        if (tag->monitor(x, h->est, tag->monitor_user)) { stopped = true; break; }
        if (h->done != VCL_RUNNING) break;
      }
      bicgstab_update_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, x, 0.0, p, 0.0, s, r, As, 0.0, Ap, r0, st, VCL_PARTIALS(b), b->tickets, &st->sums[0]);
      VCL_LAUNCHED(b, "bicgstab_update_kernel");
    }
    launched += nb;
    VCL_TRY(pull_state(b));
    if (h->done != VCL_RUNNING) break;
  }
  tag->iters = h->iters;
  tag->error = h->residual_norm / norm_rhs;                                          // bicgstab.hpp:212
  return ViennaCLSuccess;
}

// ------------------------------------------------------------------------------------------------
// Left-preconditioned BiCGStab with Jacobi  (bicgstab.hpp:398-489, jacobi_precond.hpp:103-130)
// Five kernels per iteration instead of the reference's 2 SpMV + 2 element_div + ~12 BLAS-1 launches + ~6 blocking
// scalar read-backs; the divide by diag(A) is folded into the SpMV epilogue.
// ------------------------------------------------------------------------------------------------
ViennaCLStatus bicgstab_jacobi(ViennaCLBackend b, const MatOp &A, const real *rhs, real *x, ViennaCLB200SolverTag *tag)
{
  VCL_RANGE("vcl:bicgstab_jacobi");
  VCL_REQUIRE(b, A.fmt == 0, "Jacobi needs the CSR matrix (row_info, linalg/sparse_matrix_operations.hpp:48-74)");
  const long long n = A.rows();
  VCL_TRY(vcl_ws_reserve(b, 7 * Carver::need(n)));
  Carver cv(b->ws);
  real *r = cv.take(n), *p = cv.take(n), *r0 = cv.take(n), *t0 = cv.take(n), *t1 = cv.take(n), *s = cv.take(n), *diag = cv.take(n);

  VCL_TRY(ViennaCLCUDADcsr_row_info(b, (int)n, A.csr.row_ptr, A.csr.col_idx, A.csr.values, diag, diag_precond_option(tag->precond)));
  VCL_CUDA(b, cudaMemsetAsync(x, 0, sizeof(real) * n, b->stream));
  real ss = 0.0;
  VCL_TRY(vcl_dot_host(b, n, rhs, 0, 1, rhs, 0, 1, &ss));
  const real norm_rhs = std::sqrt(ss);
  if (norm_rhs <= tag->abs_tolerance) return ViennaCLSuccess;                       // bicgstab.hpp:424-425

  SolverState *h = VCL_HSTATE(b);
  std::memset(h, 0, sizeof(SolverState));
  h->norm_rhs = norm_rhs; h->norm_rhs_sq = norm_rhs * norm_rhs; h->residual_norm = norm_rhs;
  h->tol = tag->tolerance; h->abs_tol = tag->abs_tolerance; h->maxit = tag->max_iterations;
  h->restart_every = tag->max_iterations_before_restart;
  h->need_restart = 1;
  VCL_TRY(push_state(b));
  SolverState *st = VCL_DSTATE(b);

  const int grid = vec_grid(b, n);
  const int batch = tag->monitor ? 1 : kBatch;
  while (true)
  {
    if (h->need_restart)
    {
      VCL_TRY(plain_prod(b, A, x, r));                                               // residual = A*x
      pbicg_restart_kernel<true><<<grid, VEC_THREADS, 0, b->stream>>>(n, rhs, r, p, r0, diag, st, VCL_PARTIALS(b), b->tickets);
      VCL_LAUNCHED(b, "pbicg_restart_kernel");
      h->need_restart = 0;
    }
    const int remaining = tag->max_iterations - h->iters;
    if (remaining <= 0) break;
    const int nb = std::min(batch, remaining);
    VCL_RANGE("vcl:batch");
    for (int k = 0; k < nb; ++k)
    {
      EpiFused<STEP_PBICG_ALPHA, true, true> e1 = {t0, p, r0, diag, VCL_PARTIALS(b), b->tickets, st, nullptr, nullptr, nullptr, {0.0, 0.0, 0.0}, nullptr};
      VCL_TRY(launch_prod(b, A, p, e1));
      pbicg_s_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, s, r, t0, st);
      VCL_LAUNCHED(b, "pbicg_s_kernel");
      EpiFused<STEP_PBICG_OMEGA, false, true> e2 = {t1, s, nullptr, diag, VCL_PARTIALS(b), b->tickets, st, nullptr, nullptr, nullptr, {0.0, 0.0, 0.0}, nullptr};
      VCL_TRY(launch_prod(b, A, s, e2));
      pbicg_xr_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, x, p, s, r, t1, r0, st, VCL_PARTIALS(b), b->tickets);
      VCL_LAUNCHED(b, "pbicg_xr_kernel");
      pbicg_p_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, p, r, t0, st);
      VCL_LAUNCHED(b, "pbicg_p_kernel");
    }
    VCL_TRY(pull_state(b));
    if (tag->monitor && tag->monitor(x, h->est, tag->monitor_user)) break;
    if (h->done != VCL_RUNNING) break;
  }
  tag->iters = h->iters;
  tag->error = h->residual_norm / norm_rhs;
  return ViennaCLSuccess;
}

ViennaCLStatus bicgstab_solve(ViennaCLBackend b, const MatOp &A, const real *rhs, real *x, ViennaCLB200SolverTag *tag)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, tag != nullptr, "null tag");
  VCL_TRY(check_matrix(b, A));
  tag->iters = 0; tag->error = 0.0;
  if (A.rows() == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, rhs && x, "null vector");
  VCL_CUDA(b, cudaSetDevice(b->device));
  if (diag_precond_option(tag->precond) >= 0) return bicgstab_jacobi(b, A, rhs, x, tag);
  return bicgstab_pipelined(b, A, rhs, x, tag);
}

// ------------------------------------------------------------------------------------------------
// GMRES(m), pipelined simpler-GMRES with classical Gram-Schmidt  (gmres.hpp:181-367)
// ------------------------------------------------------------------------------------------------
ViennaCLStatus gmres_solve(ViennaCLBackend b, const MatOp &A, const real *rhs, real *x, ViennaCLB200SolverTag *tag)
{
  VCL_RANGE("vcl:gmres");
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, tag != nullptr, "null tag");
  VCL_TRY(check_matrix(b, A));
  // Jacobi: the pipelined cycle runs on D^-1 A (the divide is folded into the SpMV epilogue) and -- like the reference's
  // preconditioned path (gmres.hpp:449-631) -- the estimate is tested after EVERY inner iteration: the host walks the
  // cycle's xi values, stops counting at the first iteration that meets the tolerance and truncates the update to it.
  const bool jac = diag_precond_option(tag->precond) >= 0;              // Jacobi or row scaling
  VCL_REQUIRE(b, jac || tag->precond == ViennaCLB200PrecondNone, "GMRES: unknown preconditioner id");
  VCL_REQUIRE(b, !jac || A.fmt == 0, "Jacobi needs the CSR matrix (row_info, linalg/sparse_matrix_operations.hpp:48-74)");
  VCL_REQUIRE(b, tag->krylov_dim >= 1 && tag->krylov_dim <= VCL_GMRES_MAX_KRYLOV, "krylov_dim must be in [1, 64]");
  const long long n = A.rows();
  tag->iters = 0; tag->error = 0.0;
  if (n == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, rhs && x, "null vector");
  VCL_CUDA(b, cudaSetDevice(b->device));

  const int m = tag->krylov_dim;
  const long long isz = (n + 127) / 128 * 128;                      // internal_size(): padded to 128 (forwards.h:385, gmres.hpp:192)
  const size_t small = Carver::need((size_t)m * m) + 4 * Carver::need(m);
  VCL_TRY(vcl_ws_reserve(b, (jac ? 2 : 1) * Carver::need(n) + Carver::need((size_t)isz * m) + small));
  Carver cv(b->ws);
  real *res = cv.take(n), *V = cv.take((size_t)isz * m), *R = cv.take((size_t)m * m);
  real *d_xi = cv.take(m), *d_h = cv.take(m), *d_coef = cv.take(m);
  real *diag = jac ? cv.take(n) : nullptr;
  real *d_nsq = VCL_DSCAL(b) + 8, *d_junk = VCL_DSCAL(b) + 9;

  std::vector<real> hR((size_t)m * m), xi(m), eta(m), coef(m, 0.0);

  VCL_CUDA(b, cudaMemsetAsync(x, 0, sizeof(real) * n, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(res, rhs, sizeof(real) * n, cudaMemcpyDeviceToDevice, b->stream));
  VCL_CUDA(b, cudaMemsetAsync(R, 0, sizeof(real) * m * m, b->stream));
  VCL_CUDA(b, cudaMemsetAsync(d_xi, 0, sizeof(real) * m, b->stream));
  VCL_CUDA(b, cudaMemsetAsync(d_coef, 0, sizeof(real) * m, b->stream));
  real ss = 0.0;
  VCL_TRY(vcl_dot_host(b, n, res, 0, 1, res, 0, 1, &ss));
  const real norm_rhs = std::sqrt(ss);                                           // of the UNpreconditioned rhs (gmres.hpp:478)
  real rho_0 = norm_rhs, rho = 1.0;
  if (jac)
  {
    if (norm_rhs <= tag->abs_tolerance) return ViennaCLSuccess;                  // gmres.hpp:480-481
    VCL_TRY(ViennaCLCUDADcsr_row_info(b, (int)n, A.csr.row_ptr, A.csr.col_idx, A.csr.values, diag, diag_precond_option(tag->precond)));
    VCL_TRY(ViennaCLCUDADelement_div(b, (int)n, res, 0, 1, res, 0, 1, diag, 0, 1));   // precond.apply(res), gmres.hpp:492
    VCL_TRY(vcl_dot_host(b, n, res, 0, 1, res, 0, 1, &ss));
    rho_0 = std::sqrt(ss);
  }

  unsigned max_restarts = (unsigned)tag->max_iterations / (unsigned)m;         // gmres.hpp:74-80
  if (max_restarts > 0 && max_restarts * (unsigned)m == (unsigned)tag->max_iterations) max_restarts -= 1;

  const int grid = scalar_grid(b, n);
  int coop_grid = 0;
  if (A.fmt == 0 && m <= GMRES_PERSISTENT_MAX_KRYLOV && persistent_cg_wanted(b, A.csr, n, 16))      // ahead up to 512^2, behind from 1024^2 (measured)
  {
    const int occ = vcl_occupancy(b, gmres_persistent_kernel, CSR_BLOCK_THREADS, CSR_SMEM_BYTES);
    coop_grid = std::max(1, std::min(b->sm_count * occ, std::max(A.csr.num_blocks, vcl_div_up(n, 2 * CSR_BLOCK_THREADS))));
  }
  for (unsigned restart = 0; restart <= max_restarts; ++restart)
  {
    if (restart > 0)
    {
      VCL_TRY(plain_prod(b, A, x, res));
      residual_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, res, rhs);
      VCL_LAUNCHED(b, "residual_kernel");
      if (jac) VCL_TRY(ViennaCLCUDADelement_div(b, (int)n, res, 0, 1, res, 0, 1, diag, 0, 1));
      VCL_TRY(vcl_dot_host(b, n, res, 0, 1, res, 0, 1, &ss));
      rho_0 = std::sqrt(ss);
    }
    if (rho_0 <= tag->abs_tolerance) break;                                     // gmres.hpp:227-228
    scale_residual_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, res, rho_0);
    VCL_LAUNCHED(b, "scale_residual_kernel");
    rho = 1.0;
    if (rho_0 / norm_rhs < tag->tolerance || rho_0 < tag->abs_tolerance)        // gmres.hpp:234-235 / :499-503
    {
      if (jac) tag->error = rho_0 / norm_rhs;
      break;
    }

    int k = 0;
    if (coop_grid > 0)
    {
      // small / medium CSR systems: the whole cycle in one cooperative kernel (persistent.cuh)
      CsrDev d = {A.csr.rows, (u32)A.csr.nnz, A.csr.row_ptr, A.csr.col_idx, A.csr.values, A.csr.row_blocks, A.csr.row_blocks + 1, A.csr.num_blocks};
      d.l2_mode = vcl_l2_mode(b, A.csr.nnz, A.csr.rows);
      long long nn = n, iszz = isz; int mm = m;
      real *partials = VCL_PARTIALS(b);
      const real *cres = res, *cdiag = diag;
      void *args[] = {&d, &nn, &iszz, &mm, &cres, &V, &cdiag, &R, &d_h, &d_xi, &partials};
      const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)gmres_persistent_kernel, dim3(coop_grid), dim3(CSR_BLOCK_THREADS), args,
                                                         (size_t)CSR_SMEM_BYTES, b->stream);
      if (ce == cudaErrorCooperativeLaunchTooLarge || ce == cudaErrorLaunchOutOfResources) { (void)cudaGetLastError(); coop_grid = 0; }
      else { VCL_CUDA(b, ce); VCL_LAUNCHED(b, "gmres_persistent_kernel"); k = m; }
    }
    if (coop_grid == 0)
    for (k = 0; k < m; ++k)
    {
      real *vk = V + (size_t)k * isz;
      const real *src = (k == 0) ? res : V + (size_t)(k - 1) * isz;
      if (jac)
      {
        EpiFused<STEP_NONE, false, true> e = {vk, src, nullptr, diag, VCL_PARTIALS(b), b->tickets, nullptr, d_nsq, d_junk, nullptr, {0.0, 0.0, 0.0}, nullptr};
        VCL_TRY(launch_prod(b, A, src, e));
      }
      else
      {
        EpiFused<STEP_NONE, false, false> e = {vk, src, nullptr, nullptr, VCL_PARTIALS(b), b->tickets, nullptr, d_nsq, d_junk, nullptr, {0.0, 0.0, 0.0}, nullptr};
        VCL_TRY(launch_prod(b, A, src, e));
      }
      if (k > 0)
      {
        VCL_TRY(launch_gs1(b, grid, V, n, isz, k, d_h, 1));
        gmres_gs2_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(V, n, isz, k, d_h, 1, R, m, d_nsq, VCL_PARTIALS(b), b->tickets);
        VCL_LAUNCHED(b, "gmres_gs2_kernel");
      }
      gmres_normalize_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, vk, res, R, k * m + k, d_nsq, d_xi + k, VCL_PARTIALS(b), b->tickets);
      VCL_LAUNCHED(b, "gmres_normalize_kernel");
    }

    VCL_CUDA(b, cudaMemcpyAsync(xi.data(), d_xi, sizeof(real) * m, cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaMemcpyAsync(hR.data(), R, sizeof(real) * m * m, cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));

    bool converged = false;
    const size_t kk = vcl_gmres_cycle_host<real>(tag, k, hR, xi, eta, coef, rho, rho_0, norm_rhs, jac, &converged);

    VCL_CUDA(b, cudaMemcpyAsync(d_coef, coef.data(), sizeof(real) * m, cudaMemcpyHostToDevice, b->stream));
    gmres_update_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, x, res, V, isz, d_coef, (int)kk);
    VCL_LAUNCHED(b, "gmres_update_kernel");
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));                               // coef (pageable) must outlive the copy

    tag->error = std::fabs(rho * rho_0 / norm_rhs);                              // gmres.hpp:360
    if (tag->monitor && tag->monitor(x, std::fabs(rho * rho_0 / norm_rhs), tag->monitor_user)) break;
    if (converged) break;                                                        // gmres.hpp:627-628
  }
  return ViennaCLSuccess;
}

// sums `chunk` entries starting at in[0] into out[0] (per-op API: tolerate producers that spread partials over a chunk)
__global__ void chunk_sum_kernel(const real *in, int chunk, real *out)
{
  __shared__ real s_red[32];
  real acc[1] = {0.0};
  for (int i = threadIdx.x; i < chunk; i += blockDim.x) acc[0] += in[i];
  block_sum<1>(acc, s_red);
  if (threadIdx.x == 0) out[0] = acc[0];
}

// Per-op API: a producer writes the FULLY REDUCED value into element 0 of its chunk of the caller's inner_prod_buffer; the
// consumers (and the reference's drivers, cg.hpp:168-172) sum whole chunks.  The rest of the chunk is zeroed here so that a
// buffer that still holds stale values -- e.g. reference-style per-block partials -- cannot leak into alpha, ||v_k||^2 or h_j.
ViennaCLStatus zero_chunk_tail(ViennaCLBackend b, real *chunk_begin, long long chunk)
{
  if (chunk > 1) VCL_CUDA(b, cudaMemsetAsync(chunk_begin + 1, 0, sizeof(real) * (size_t)(chunk - 1), b->stream));
  return ViennaCLSuccess;
}

MatOp from_csr(const ViennaCLCUDADcsr *A) { MatOp m; m.fmt = 0; m.csr = *A; m.sell = ViennaCLCUDADsell(); m.hyb = ViennaCLCUDADhyb(); return m; }
MatOp from_sell(const ViennaCLCUDADsell *A) { MatOp m; m.fmt = 1; m.sell = *A; m.csr = ViennaCLCUDADcsr(); m.hyb = ViennaCLCUDADhyb(); return m; }
MatOp from_hyb(const ViennaCLCUDADhyb *A) { MatOp m; m.fmt = 2; m.hyb = *A; m.csr = ViennaCLCUDADcsr(); m.sell = ViennaCLCUDADsell(); return m; }
MatOp from_ell(const ViennaCLCUDADell *A) { ViennaCLCUDADhyb h = ViennaCLCUDADhyb(); h.ell = *A; return from_hyb(&h); }

ViennaCLStatus fused_prod_api(ViennaCLBackend b, const MatOp &A, const real *p, real *Ap, const real *r0,
                              real *out_ApAp, real *out_pAp, real *out_Apr0, long long chunk)
{
  VCL_CHECK_BACKEND(b);
  VCL_TRY(check_matrix(b, A));
  if (A.rows() == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, p && Ap && p != Ap, "bad vectors");
  VCL_TRY(zero_chunk_tail(b, out_ApAp, chunk));
  VCL_TRY(zero_chunk_tail(b, out_pAp, chunk));
  if (r0 && out_Apr0 && out_Apr0 != out_ApAp && out_Apr0 != out_pAp) VCL_TRY(zero_chunk_tail(b, out_Apr0, chunk));
  if (r0)
  {
    EpiFused<STEP_NONE, true, false> e = {Ap, p, r0, nullptr, VCL_PARTIALS(b), b->tickets, nullptr, out_ApAp, out_pAp, out_Apr0, {0.0, 0.0, 0.0}, nullptr};
    return launch_prod(b, A, p, e);
  }
  EpiFused<STEP_NONE, false, false> e = {Ap, p, nullptr, nullptr, VCL_PARTIALS(b), b->tickets, nullptr, out_ApAp, out_pAp, nullptr, {0.0, 0.0, 0.0}, nullptr};
  return launch_prod(b, A, p, e);
}

} // namespace

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

ViennaCLStatus ViennaCLCUDADcsr_cg(ViennaCLBackend b, const ViennaCLCUDADcsr *A, const real *rhs, real *x, ViennaCLB200SolverTag *tag)
{ VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A, "null matrix"); return cg_solve(b, from_csr(A), rhs, x, tag); }
ViennaCLStatus ViennaCLCUDADsell_cg(ViennaCLBackend b, const ViennaCLCUDADsell *A, const real *rhs, real *x, ViennaCLB200SolverTag *tag)
{ VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A, "null matrix"); return cg_solve(b, from_sell(A), rhs, x, tag); }
ViennaCLStatus ViennaCLCUDADcsr_bicgstab(ViennaCLBackend b, const ViennaCLCUDADcsr *A, const real *rhs, real *x, ViennaCLB200SolverTag *tag)
{ VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A, "null matrix"); return bicgstab_solve(b, from_csr(A), rhs, x, tag); }
ViennaCLStatus ViennaCLCUDADsell_bicgstab(ViennaCLBackend b, const ViennaCLCUDADsell *A, const real *rhs, real *x, ViennaCLB200SolverTag *tag)
{ VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A, "null matrix"); return bicgstab_solve(b, from_sell(A), rhs, x, tag); }
ViennaCLStatus ViennaCLCUDADcsr_gmres(ViennaCLBackend b, const ViennaCLCUDADcsr *A, const real *rhs, real *x, ViennaCLB200SolverTag *tag)
{ VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A, "null matrix"); return gmres_solve(b, from_csr(A), rhs, x, tag); }
ViennaCLStatus ViennaCLCUDADsell_gmres(ViennaCLBackend b, const ViennaCLCUDADsell *A, const real *rhs, real *x, ViennaCLB200SolverTag *tag)
{ VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A, "null matrix"); return gmres_solve(b, from_sell(A), rhs, x, tag); }

#define VCL_SOLVER_ENTRY(name, type, conv, fn) \
ViennaCLStatus name(ViennaCLBackend b, const type *A, const real *rhs, real *x, ViennaCLB200SolverTag *tag) \
{ VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A, "null matrix"); return fn(b, conv(A), rhs, x, tag); }
VCL_SOLVER_ENTRY(ViennaCLCUDADell_cg, ViennaCLCUDADell, from_ell, cg_solve)
VCL_SOLVER_ENTRY(ViennaCLCUDADhyb_cg, ViennaCLCUDADhyb, from_hyb, cg_solve)
VCL_SOLVER_ENTRY(ViennaCLCUDADell_bicgstab, ViennaCLCUDADell, from_ell, bicgstab_solve)
VCL_SOLVER_ENTRY(ViennaCLCUDADhyb_bicgstab, ViennaCLCUDADhyb, from_hyb, bicgstab_solve)
VCL_SOLVER_ENTRY(ViennaCLCUDADell_gmres, ViennaCLCUDADell, from_ell, gmres_solve)
VCL_SOLVER_ENTRY(ViennaCLCUDADhyb_gmres, ViennaCLCUDADhyb, from_hyb, gmres_solve)

// plain products for the other formats (linalg/sparse_matrix_operations.hpp:90-121 dispatch)
static ViennaCLStatus ellhyb_mv(ViennaCLBackend b, const MatOp &A, const real *x, ViennaCLInt offx, ViennaCLInt incx, real alpha,
                                real *y, ViennaCLInt offy, ViennaCLInt incy, real beta)
{
  VCL_TRY(check_matrix_any(b, A));
  if (A.rows() == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, x && y && incx != 0 && incy != 0 && x != y, "bad vectors");
  EpiAxpby epi = {y, offy, incy, alpha, beta};
  return vcl_launch_ell(b, A.hyb, make_xvec(x, offx, incx), epi);
}
ViennaCLStatus ViennaCLCUDADellmv(ViennaCLBackend b, const ViennaCLCUDADell *A, const real *x, ViennaCLInt offx, ViennaCLInt incx, real alpha,
                                  real *y, ViennaCLInt offy, ViennaCLInt incy, real beta)
{ VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A, "null matrix"); return ellhyb_mv(b, from_ell(A), x, offx, incx, alpha, y, offy, incy, beta); }
ViennaCLStatus ViennaCLCUDADhybmv(ViennaCLBackend b, const ViennaCLCUDADhyb *A, const real *x, ViennaCLInt offx, ViennaCLInt incx, real alpha,
                                  real *y, ViennaCLInt offy, ViennaCLInt incy, real beta)
{ VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A, "null matrix"); return ellhyb_mv(b, from_hyb(A), x, offx, incx, alpha, y, offy, incy, beta); }

ViennaCLStatus ViennaCLCUDADpipelined_cg_prod_ell(ViennaCLBackend b, const ViennaCLCUDADell *A, const real *p, real *Ap, real *buf, ViennaCLInt buf_size)
{ VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A && buf && buf_size >= 3, "bad arguments"); const int chunk = buf_size / 3;
  return fused_prod_api(b, from_ell(A), p, Ap, nullptr, buf + chunk, buf + 2 * chunk, nullptr, chunk); }
ViennaCLStatus ViennaCLCUDADpipelined_cg_prod_hyb(ViennaCLBackend b, const ViennaCLCUDADhyb *A, const real *p, real *Ap, real *buf, ViennaCLInt buf_size)
{ VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A && buf && buf_size >= 3, "bad arguments"); const int chunk = buf_size / 3;
  return fused_prod_api(b, from_hyb(A), p, Ap, nullptr, buf + chunk, buf + 2 * chunk, nullptr, chunk); }
ViennaCLStatus ViennaCLCUDADpipelined_bicgstab_prod_ell(ViennaCLBackend b, const ViennaCLCUDADell *A, const real *p, real *Ap,
                                                        const real *r0star, real *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset)
{ VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A && buf && r0star && chunk >= 1, "bad arguments");
  return fused_prod_api(b, from_ell(A), p, Ap, r0star, buf + chunk, buf + 2 * (size_t)chunk, buf + chunk_offset, chunk); }
ViennaCLStatus ViennaCLCUDADpipelined_bicgstab_prod_hyb(ViennaCLBackend b, const ViennaCLCUDADhyb *A, const real *p, real *Ap,
                                                        const real *r0star, real *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset)
{ VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A && buf && r0star && chunk >= 1, "bad arguments");
  return fused_prod_api(b, from_hyb(A), p, Ap, r0star, buf + chunk, buf + 2 * (size_t)chunk, buf + chunk_offset, chunk); }
ViennaCLStatus ViennaCLCUDADpipelined_gmres_prod_ell(ViennaCLBackend b, const ViennaCLCUDADell *A, const real *p, real *Ap, real *buf, ViennaCLInt buf_size)
{ return ViennaCLCUDADpipelined_cg_prod_ell(b, A, p, Ap, buf, buf_size); }
ViennaCLStatus ViennaCLCUDADpipelined_gmres_prod_hyb(ViennaCLBackend b, const ViennaCLCUDADhyb *A, const real *p, real *Ap, real *buf, ViennaCLInt buf_size)
{ return ViennaCLCUDADpipelined_cg_prod_hyb(b, A, p, Ap, buf, buf_size); }

// ---- per-step entry points (linalg/iterative_operations.hpp) ----
ViennaCLStatus ViennaCLCUDADpipelined_cg_vector_update(ViennaCLBackend b, ViennaCLInt n, real *result, real alpha,
                                                       real *p, real *r, const real *Ap, real beta,
                                                       real *buf, ViennaCLInt buf_size)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, n >= 0 && buf && buf_size >= 3, "bad arguments");
  if (n == 0) return ViennaCLSuccess;
  VCL_TRY(zero_chunk_tail(b, buf, buf_size / 3));
  cg_update_kernel<<<vec_grid(b, n), VEC_THREADS, 0, b->stream>>>(n, result, p, r, Ap, alpha, beta, nullptr, VCL_PARTIALS(b), b->tickets, buf);
  VCL_LAUNCHED(b, "cg_update_kernel");
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDADpipelined_cg_prod_csr(ViennaCLBackend b, const ViennaCLCUDADcsr *A, const real *p, real *Ap,
                                                  real *buf, ViennaCLInt buf_size)
{
  VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A && buf && buf_size >= 3, "bad arguments");
  const int chunk = buf_size / 3;
  return fused_prod_api(b, from_csr(A), p, Ap, nullptr, buf + chunk, buf + 2 * chunk, nullptr, chunk);
}

ViennaCLStatus ViennaCLCUDADpipelined_cg_prod_sell(ViennaCLBackend b, const ViennaCLCUDADsell *A, const real *p, real *Ap,
                                                   real *buf, ViennaCLInt buf_size)
{
  VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A && buf && buf_size >= 3, "bad arguments");
  const int chunk = buf_size / 3;
  return fused_prod_api(b, from_sell(A), p, Ap, nullptr, buf + chunk, buf + 2 * chunk, nullptr, chunk);
}

ViennaCLStatus ViennaCLCUDADpipelined_bicgstab_update_s(ViennaCLBackend b, ViennaCLInt n, real *s, const real *r, const real *Ap,
                                                        real *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, n >= 0 && buf && chunk >= 1, "bad arguments");
  if (n == 0) return ViennaCLSuccess;
  chunk_sum_kernel<<<1, 256, 0, b->stream>>>(buf, chunk, VCL_DSCAL(b) + 16);
  VCL_LAUNCHED(b, "chunk_sum_kernel");
  chunk_sum_kernel<<<1, 256, 0, b->stream>>>(buf + 3 * (size_t)chunk, chunk, VCL_DSCAL(b) + 17);
  VCL_LAUNCHED(b, "chunk_sum_kernel");
  VCL_TRY(zero_chunk_tail(b, buf + chunk_offset, chunk));
  bicgstab_update_s_kernel<<<vec_grid(b, n), VEC_THREADS, 0, b->stream>>>(n, s, r, Ap, VCL_DSCAL(b) + 16, VCL_DSCAL(b) + 17, nullptr,
                                                                        VCL_PARTIALS(b), b->tickets, buf + chunk_offset);
  VCL_LAUNCHED(b, "bicgstab_update_s_kernel");
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDADpipelined_bicgstab_vector_update(ViennaCLBackend b, ViennaCLInt n, real *result, real alpha, real *p,
                                                             real omega, const real *s, real *residual, const real *As,
                                                             real beta, const real *Ap, const real *r0star,
                                                             real *buf, ViennaCLInt chunk)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, n >= 0 && buf, "bad arguments");
  if (n == 0) return ViennaCLSuccess;
  VCL_TRY(zero_chunk_tail(b, buf, chunk));
  bicgstab_update_kernel<<<vec_grid(b, n), VEC_THREADS, 0, b->stream>>>(n, result, alpha, p, omega, s, residual, As, beta, Ap, r0star,
                                                                      nullptr, VCL_PARTIALS(b), b->tickets, buf);
  VCL_LAUNCHED(b, "bicgstab_update_kernel");
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDADpipelined_bicgstab_prod_csr(ViennaCLBackend b, const ViennaCLCUDADcsr *A, const real *p, real *Ap,
                                                        const real *r0star, real *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset)
{
  VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A && buf && r0star && chunk >= 1, "bad arguments");
  return fused_prod_api(b, from_csr(A), p, Ap, r0star, buf + chunk, buf + 2 * (size_t)chunk, buf + chunk_offset, chunk);
}

ViennaCLStatus ViennaCLCUDADpipelined_bicgstab_prod_sell(ViennaCLBackend b, const ViennaCLCUDADsell *A, const real *p, real *Ap,
                                                         const real *r0star, real *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset)
{
  VCL_CHECK_BACKEND(b); VCL_REQUIRE(b, A && buf && r0star && chunk >= 1, "bad arguments");
  return fused_prod_api(b, from_sell(A), p, Ap, r0star, buf + chunk, buf + 2 * (size_t)chunk, buf + chunk_offset, chunk);
}

ViennaCLStatus ViennaCLCUDADpipelined_gmres_normalize_vk(ViennaCLBackend b, ViennaCLInt n, real *v_k, const real *residual,
                                                         real *R, ViennaCLInt offset_in_R, const real *buf,
                                                         real *r_dot_vk, ViennaCLInt chunk, ViennaCLInt chunk_offset)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, n >= 0 && v_k && residual && R && buf && r_dot_vk && chunk >= 1, "bad arguments");
  if (n == 0) return ViennaCLSuccess;
  chunk_sum_kernel<<<1, 256, 0, b->stream>>>(buf + chunk, chunk, VCL_DSCAL(b) + 16);      // ||v_k||^2 lives in chunk 1
  VCL_LAUNCHED(b, "chunk_sum_kernel");
  VCL_TRY(zero_chunk_tail(b, r_dot_vk + chunk_offset, chunk));
  gmres_normalize_kernel<<<scalar_grid(b, n), VEC_THREADS, 0, b->stream>>>(n, v_k, residual, R, offset_in_R, VCL_DSCAL(b) + 16,
                                                                         r_dot_vk + chunk_offset, VCL_PARTIALS(b), b->tickets);
  VCL_LAUNCHED(b, "gmres_normalize_kernel");
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDADpipelined_gmres_gram_schmidt_stage1(ViennaCLBackend b, const real *basis, ViennaCLInt n,
                                                                ViennaCLInt internal_n, ViennaCLInt k, real *vi_in_vk, ViennaCLInt chunk)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, n >= 0 && basis && vi_in_vk && k >= 0 && k < VCL_GMRES_MAX_KRYLOV && chunk >= 1, "bad arguments");
  if (n == 0 || k == 0) return ViennaCLSuccess;
  if (chunk > 1) VCL_CUDA(b, cudaMemsetAsync(vi_in_vk, 0, sizeof(real) * (size_t)k * (size_t)chunk, b->stream));
  return launch_gs1(b, scalar_grid(b, n), basis, n, internal_n, k, vi_in_vk, chunk);
}

ViennaCLStatus ViennaCLCUDADpipelined_gmres_gram_schmidt_stage2(ViennaCLBackend b, real *basis, ViennaCLInt n,
                                                                ViennaCLInt internal_n, ViennaCLInt k, const real *vi_in_vk,
                                                                real *R, ViennaCLInt krylov_dim, real *buf, ViennaCLInt chunk)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, n >= 0 && basis && vi_in_vk && R && buf && k >= 0 && k < VCL_GMRES_MAX_KRYLOV && chunk >= 1, "bad arguments");
  if (n == 0) return ViennaCLSuccess;
  // second reduction stage of <v_i, v_k>: fold each chunk into its first element (no-op for our own stage 1)
  for (int j = 0; j < k; ++j)
  {
    chunk_sum_kernel<<<1, 256, 0, b->stream>>>(vi_in_vk + (size_t)j * chunk, chunk, VCL_DSCAL(b) + VCL_DSCAL_GS_FOLD + j);
    VCL_LAUNCHED(b, "chunk_sum_kernel");
  }
  VCL_REQUIRE(b, (internal_n & 1) == 0 && (reinterpret_cast<uintptr_t>(basis) & 15u) == 0u, "Krylov basis must be 16-byte aligned with an even internal size");
  VCL_TRY(zero_chunk_tail(b, buf + chunk, chunk));
  gmres_gs2_kernel<<<scalar_grid(b, n), VEC_THREADS, 0, b->stream>>>(basis, n, internal_n, k, VCL_DSCAL(b) + VCL_DSCAL_GS_FOLD, 1, R, krylov_dim,
                                                                   buf + chunk, VCL_PARTIALS(b), b->tickets);
  VCL_LAUNCHED(b, "gmres_gs2_kernel");
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDADpipelined_gmres_update_result(ViennaCLBackend b, ViennaCLInt n, real *result, const real *residual,
                                                          const real *basis, ViennaCLInt internal_n, const real *coefficients, ViennaCLInt k)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, n >= 0 && result && residual && coefficients && k >= 0 && k <= VCL_GMRES_MAX_KRYLOV, "bad arguments");
  if (n == 0) return ViennaCLSuccess;
  gmres_update_kernel<<<scalar_grid(b, n), VEC_THREADS, 0, b->stream>>>(n, result, residual, basis, internal_n, coefficients, k);
  VCL_LAUNCHED(b, "gmres_update_kernel");
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDADpipelined_gmres_prod_csr(ViennaCLBackend b, const ViennaCLCUDADcsr *A, const real *p, real *Ap,
                                                     real *buf, ViennaCLInt buf_size)
{ return ViennaCLCUDADpipelined_cg_prod_csr(b, A, p, Ap, buf, buf_size); }

ViennaCLStatus ViennaCLCUDADpipelined_gmres_prod_sell(ViennaCLBackend b, const ViennaCLCUDADsell *A, const real *p, real *Ap,
                                                      real *buf, ViennaCLInt buf_size)
{ return ViennaCLCUDADpipelined_cg_prod_sell(b, A, p, Ap, buf, buf_size); }

} // extern "C"
} // namespace VCL_NS
