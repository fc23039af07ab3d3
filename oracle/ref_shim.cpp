// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" shim over the UNMODIFIED reference headers under /root/reference, compiled
// with its OpenMP host backend (-DVIENNACL_WITH_OPENMP).  The resulting shared object lives in
// oracle/_ref/ (git-ignored, travels to the GPU box) and is used
//   * to pin oracle/vcl_oracle.c (the plain-C restatement) and to generate tests/golden/*,
//   * as the "reference" CPU baseline timed by bench.py (cpu_baseline leg / --impl reference).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// Reference entry points exercised (all relative to /root/reference):
//   viennacl/linalg/sparse_matrix_operations.hpp:90-121   prod_impl dispatch (CSR / SELL)
//   viennacl/linalg/host_based/sparse_matrix_operations.hpp:110-186, 1796-1858, 52-98
//   viennacl/linalg/cg.hpp:128-187          pipelined CG
//   viennacl/linalg/bicgstab.hpp:97-215     pipelined BiCGStab; :398-489 preconditioned BiCGStab
//   viennacl/linalg/gmres.hpp:181-367       pipelined GMRES;    :449-631 Householder GMRES
//   viennacl/linalg/jacobi_precond.hpp:103-130
//   viennacl/ell_matrix.hpp:122-166, viennacl/hyb_matrix.hpp:127-214 (host -> ELL / HYB layout),
//   viennacl/linalg/host_based/sparse_matrix_operations.hpp:1503-1538 (ELL), :1873-1927 (HYB)
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/sliced_ell_matrix.hpp"
#include "viennacl/ell_matrix.hpp"
#include "viennacl/hyb_matrix.hpp"
#include "viennacl/coordinate_matrix.hpp"
#include "viennacl/linalg/prod.hpp"
#include "viennacl/linalg/inner_prod.hpp"
#include "viennacl/linalg/norm_2.hpp"
#include "viennacl/linalg/cg.hpp"
#include "viennacl/linalg/bicgstab.hpp"
#include "viennacl/linalg/gmres.hpp"
#include "viennacl/linalg/jacobi_precond.hpp"
#ifndef VCLREF_F32
#include "viennacl/linalg/mixed_precision_cg.hpp"
#endif

typedef unsigned int u32;
// element type: built twice, as it is (double -> libvcl_ref.so) and with -DVCLREF_F32 (float -> libvcl_ref_f32.so)
#ifdef VCLREF_F32
typedef float real_t;
#else
typedef double real_t;
#endif
typedef viennacl::compressed_matrix<real_t> csr_t;
typedef viennacl::sliced_ell_matrix<real_t> sell_t;
typedef viennacl::vector<real_t> vec_t;
typedef viennacl::ell_matrix<real_t> ell_t;
typedef viennacl::hyb_matrix<real_t> hyb_t;

namespace {

// Read-only iterator view over raw CSR arrays so that the reference's generic
// copy(CPUMatrixT, sliced_ell_matrix) (sliced_ell_matrix.hpp:140-214) can build SELL without a
// std::vector<std::map<>> detour.
struct raw_csr_view
{
  typedef std::size_t size_type; typedef real_t value_type;
  std::size_t rows_, cols_; const u32 *rp_, *ci_; const real_t *v_;
  struct const_iterator2
  {
    const raw_csr_view *m; std::size_t row, k;
    std::size_t index1() const { return row; }
    std::size_t index2() const { return m->ci_[k]; }
    real_t operator*() const { return m->v_[k]; }
    const_iterator2 & operator++() { ++k; return *this; }
    bool operator!=(const_iterator2 const & o) const { return k != o.k; }
    bool operator==(const_iterator2 const & o) const { return k == o.k; }
  };
  struct const_iterator1
  {
    const raw_csr_view *m; std::size_t row;
    std::size_t index1() const { return row; }
    const_iterator1 & operator++() { ++row; return *this; }
    bool operator!=(const_iterator1 const & o) const { return row != o.row; }
    bool operator==(const_iterator1 const & o) const { return row == o.row; }
    const_iterator2 begin() const { const_iterator2 it = {m, row, m->rp_[row]}; return it; }
    const_iterator2 end()   const { const_iterator2 it = {m, row, m->rp_[row+1]}; return it; }
  };
  std::size_t size1() const { return rows_; }
  std::size_t size2() const { return cols_; }
  const_iterator1 begin1() const { const_iterator1 it = {this, 0}; return it; }
  const_iterator1 end1()   const { const_iterator1 it = {this, rows_}; return it; }
};

struct history
{
  double *buf; int cap; int len;
};

bool monitor_cb(vec_t const &, real_t est, void *user)
{
  history *h = static_cast<history*>(user);
  if (h->buf && h->len < h->cap) h->buf[h->len] = est;
  h->len++;
  return false;
}

// A do-nothing preconditioner: routes solve() into the reference's *generic* code paths
// (bicgstab.hpp:398-489, gmres.hpp:449-631) without changing the mathematics.
struct identity_precond
{
  template<typename V> void apply(V &) const {}
};

template<typename MatT>
int run_solver(MatT const & A, int solver, int precond, csr_t const * A_csr_for_jacobi,
               vec_t const & b, vec_t & x, double tol, double abs_tol, int maxit, int krylov, int restart_every,
               int *iters, double *err, history *h)
{
  bool (*mon)(vec_t const &, real_t, void*) = h ? monitor_cb : NULL;
  if (solver == 0)
  {
    viennacl::linalg::cg_tag tag(tol, maxit); tag.abs_tolerance(abs_tol);
    viennacl::linalg::cg_solver<vec_t> s(tag);
    if (mon) s.set_monitor(mon, h);
    if (precond == 0) x = s(A, b);
    else if (precond == 2) x = s(A, b, identity_precond());
    else if (precond == 1)
    {
      if (!A_csr_for_jacobi) return 2;
      viennacl::linalg::jacobi_precond<csr_t> jac(*A_csr_for_jacobi, viennacl::linalg::jacobi_tag());
      x = s(A, b, jac);                          // generic PCG, cg.hpp:257-322
    }
    else return 2;
    *iters = int(s.tag().iters()); *err = s.tag().error();
    return 0;
  }
  if (solver == 1)
  {
    viennacl::linalg::bicgstab_tag tag(tol, maxit, restart_every > 0 ? restart_every : 200); tag.abs_tolerance(abs_tol);
    viennacl::linalg::bicgstab_solver<vec_t> s(tag);
    if (mon) s.set_monitor(mon, h);
    if (precond == 0) x = s(A, b);
    else if (precond == 1)
    {
      if (!A_csr_for_jacobi) return 2;
      viennacl::linalg::jacobi_precond<csr_t> jac(*A_csr_for_jacobi, viennacl::linalg::jacobi_tag());
      x = s(A, b, jac);
    }
    else x = s(A, b, identity_precond());
    *iters = int(s.tag().iters()); *err = s.tag().error();
    return 0;
  }
  if (solver == 2)
  {
    viennacl::linalg::gmres_tag tag(tol, maxit, krylov); tag.abs_tolerance(abs_tol);
    viennacl::linalg::gmres_solver<vec_t> s(tag);
    if (mon) s.set_monitor(mon, h);
    if (precond == 0) x = s(A, b);               // pipelined (host variant is defective)
    else if (precond == 1)
    {
      if (!A_csr_for_jacobi) return 2;
      viennacl::linalg::jacobi_precond<csr_t> jac(*A_csr_for_jacobi, viennacl::linalg::jacobi_tag());
      x = s(A, b, jac);
    }
    else x = s(A, b, identity_precond());        // Householder path, the usable GMRES oracle
    *iters = int(s.tag().iters()); *err = s.tag().error();
    return 0;
  }
  return 1;
}

} // namespace

extern "C" {

int vclref_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void vclref_set_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// y[offy + i*incy] = alpha * (A x)_i + beta * y_i     (x read at offx + col*incx)
// mode: 0 = prod_impl(A,x,alpha,y,beta); 1 = y = prod(A,x); 2 = y += prod(A,x); 3 = y -= prod(A,x); 4 = x = prod(A,x) (aliasing, result in x)
int vclref_csr_spmv(int rows, int cols, int nnz, const u32 *rp, const u32 *ci, const real_t *v,
                    real_t *x, int offx, int incx, int nx,
                    real_t alpha,
                    real_t *y, int offy, int incy, int ny,
                    real_t beta, int mode)
{
  csr_t A(const_cast<u32*>(rp), const_cast<u32*>(ci), const_cast<real_t*>(v), viennacl::MAIN_MEMORY, rows, cols, nnz);
  vec_t vx(x, viennacl::MAIN_MEMORY, std::size_t(nx), std::size_t(offx), std::size_t(incx));
  vec_t vy(y, viennacl::MAIN_MEMORY, std::size_t(ny), std::size_t(offy), std::size_t(incy));
  switch (mode)
  {
  case 0: viennacl::linalg::prod_impl(A, vx, alpha, vy, beta); break;
  case 1: vy  = viennacl::linalg::prod(A, vx); break;
  case 2: vy += viennacl::linalg::prod(A, vx); break;
  case 3: vy -= viennacl::linalg::prod(A, vx); break;
  case 4: vx  = viennacl::linalg::prod(A, vx); break;
  default: return 1;
  }
  return 0;
}

// Builds SELL-C (sigma = 1) from CSR through the reference's own copy(); arrays are malloc'ed here, freed by vclref_free.
int vclref_sell_build(int rows, int cols, const u32 *rp, const u32 *ci, const real_t *v, int C,
                      u32 **cols_per_block, u32 **block_start, u32 **col_idx, real_t **elements,
                      int *num_blocks, long long *padded_nnz)
{
  raw_csr_view view = {std::size_t(rows), std::size_t(cols), rp, ci, v};
  sell_t S((std::size_t(rows)), (std::size_t(cols)), (std::size_t(C)));
  viennacl::copy(view, S);
  std::size_t nb = (std::size_t(rows) - 1) / S.rows_per_block() + 1;
  std::size_t tot = S.handle().raw_size() / sizeof(real_t);
  *num_blocks = int(nb); *padded_nnz = (long long)tot;
  *cols_per_block = (u32*)std::malloc(sizeof(u32) * nb);
  *block_start    = (u32*)std::malloc(sizeof(u32) * nb);
  *col_idx        = (u32*)std::malloc(sizeof(u32) * (tot ? tot : 1));
  *elements       = (real_t*)std::malloc(sizeof(real_t) * (tot ? tot : 1));
  std::memcpy(*cols_per_block, S.handle1().ram_handle().get(), sizeof(u32) * nb);
  std::memcpy(*block_start,    S.handle3().ram_handle().get(), sizeof(u32) * nb);
  std::memcpy(*col_idx,        S.handle2().ram_handle().get(), sizeof(u32) * tot);
  std::memcpy(*elements,       S.handle().ram_handle().get(),  sizeof(real_t) * tot);
  return 0;
}

void vclref_free(void *p) { std::free(p); }

// Reference host SELL SpMV.  Refuses rows % C == 0 because the reference over-reads its block arrays
// there (host_based/sparse_matrix_operations.hpp:1810 vs sliced_ell_matrix.hpp:154); see SURVEY 8c-2.
int vclref_sell_spmv(int rows, int cols, const u32 *rp, const u32 *ci, const real_t *v, int C,
                     real_t *x, real_t alpha, real_t *y, real_t beta)
{
  if (rows % C == 0) return 3;
  raw_csr_view view = {std::size_t(rows), std::size_t(cols), rp, ci, v};
  sell_t S((std::size_t(rows)), (std::size_t(cols)), (std::size_t(C)));
  viennacl::copy(view, S);
  vec_t vx(x, viennacl::MAIN_MEMORY, std::size_t(cols));
  vec_t vy(y, viennacl::MAIN_MEMORY, std::size_t(rows));
  viennacl::linalg::prod_impl(S, vx, alpha, vy, beta);
  return 0;
}

// ELL / HYB built by the reference's own copy(); arrays malloc'ed here (vclref_free).  ELL: coords/elements hold
// internal_size1 * internal_maxnnz entries, entry j of row r at j*internal_size1 + r.  HYB: ELL part of width ell_width
// (csr_threshold 0.8) + CSR remainder.
int vclref_ell_build(int rows, int cols, const u32 *rp, const u32 *ci, const real_t *v,
                     u32 **coords, real_t **elements, int *maxnnz, int *internal_rows)
{
  raw_csr_view view = {std::size_t(rows), std::size_t(cols), rp, ci, v};
  ell_t E(viennacl::context(viennacl::MAIN_MEMORY));
  viennacl::copy(view, E);
  std::size_t tot = E.internal_nnz();
  *maxnnz = int(E.internal_maxnnz()); *internal_rows = int(E.internal_size1());
  *coords = (u32*)std::malloc(sizeof(u32) * (tot ? tot : 1));
  *elements = (real_t*)std::malloc(sizeof(real_t) * (tot ? tot : 1));
  std::memcpy(*coords, E.handle2().ram_handle().get(), sizeof(u32) * tot);
  std::memcpy(*elements, E.handle().ram_handle().get(), sizeof(real_t) * tot);
  return 0;
}

int vclref_ell_spmv(int rows, int cols, const u32 *rp, const u32 *ci, const real_t *v,
                    real_t *x, real_t alpha, real_t *y, real_t beta)
{
  raw_csr_view view = {std::size_t(rows), std::size_t(cols), rp, ci, v};
  ell_t E(viennacl::context(viennacl::MAIN_MEMORY));
  viennacl::copy(view, E);
  vec_t vx(x, viennacl::MAIN_MEMORY, std::size_t(cols));
  vec_t vy(y, viennacl::MAIN_MEMORY, std::size_t(rows));
  viennacl::linalg::prod_impl(E, vx, alpha, vy, beta);
  return 0;
}

int vclref_hyb_build(int rows, int cols, const u32 *rp, const u32 *ci, const real_t *v,
                     u32 **ell_coords, real_t **ell_elements, int *ell_width, int *internal_rows,
                     u32 **csr_rows, u32 **csr_cols, real_t **csr_elements, int *csr_nnz)
{
  raw_csr_view view = {std::size_t(rows), std::size_t(cols), rp, ci, v};
  hyb_t H(viennacl::context(viennacl::MAIN_MEMORY));
  viennacl::copy(view, H);
  std::size_t tot = H.internal_size1() * H.internal_ellnnz();
  *ell_width = int(H.internal_ellnnz()); *internal_rows = int(H.internal_size1()); *csr_nnz = int(H.csr_nnz());
  *ell_coords = (u32*)std::malloc(sizeof(u32) * (tot ? tot : 1));
  *ell_elements = (real_t*)std::malloc(sizeof(real_t) * (tot ? tot : 1));
  *csr_rows = (u32*)std::malloc(sizeof(u32) * (std::size_t(rows) + 1));
  *csr_cols = (u32*)std::malloc(sizeof(u32) * H.csr_nnz());
  *csr_elements = (real_t*)std::malloc(sizeof(real_t) * H.csr_nnz());
  std::memcpy(*ell_coords, H.handle2().ram_handle().get(), sizeof(u32) * tot);
  std::memcpy(*ell_elements, H.handle().ram_handle().get(), sizeof(real_t) * tot);
  std::memcpy(*csr_rows, H.handle3().ram_handle().get(), sizeof(u32) * (std::size_t(rows) + 1));
  std::memcpy(*csr_cols, H.handle4().ram_handle().get(), sizeof(u32) * H.csr_nnz());
  std::memcpy(*csr_elements, H.handle5().ram_handle().get(), sizeof(real_t) * H.csr_nnz());
  return 0;
}

int vclref_hyb_spmv(int rows, int cols, const u32 *rp, const u32 *ci, const real_t *v,
                    real_t *x, real_t alpha, real_t *y, real_t beta)
{
  raw_csr_view view = {std::size_t(rows), std::size_t(cols), rp, ci, v};
  hyb_t H(viennacl::context(viennacl::MAIN_MEMORY));
  viennacl::copy(view, H);
  vec_t vx(x, viennacl::MAIN_MEMORY, std::size_t(cols));
  vec_t vy(y, viennacl::MAIN_MEMORY, std::size_t(rows));
  viennacl::linalg::prod_impl(H, vx, alpha, vy, beta);
  return 0;
}

// COO built by the reference's copy() (coordinate_matrix.hpp:47-102): coords = (row, col) pairs; product host_based/...:1222-1247
int vclref_coo_build(int rows, int cols, const u32 *rp, const u32 *ci, const real_t *v, u32 **coords, real_t **elements, int *nnz)
{
  raw_csr_view view = {std::size_t(rows), std::size_t(cols), rp, ci, v};
  viennacl::coordinate_matrix<real_t> M(viennacl::context(viennacl::MAIN_MEMORY));
  viennacl::copy(view, M);
  std::size_t n = M.nnz();
  *nnz = int(n);
  *coords = (u32*)std::malloc(sizeof(u32) * 2 * (n ? n : 1));
  *elements = (real_t*)std::malloc(sizeof(real_t) * (n ? n : 1));
  std::memcpy(*coords, M.handle12().ram_handle().get(), sizeof(u32) * 2 * n);
  std::memcpy(*elements, M.handle().ram_handle().get(), sizeof(real_t) * n);
  return 0;
}

int vclref_coo_spmv(int rows, int cols, const u32 *rp, const u32 *ci, const real_t *v,
                    real_t *x, real_t alpha, real_t *y, real_t beta)
{
  raw_csr_view view = {std::size_t(rows), std::size_t(cols), rp, ci, v};
  viennacl::coordinate_matrix<real_t> M(viennacl::context(viennacl::MAIN_MEMORY));
  viennacl::copy(view, M);
  vec_t vx(x, viennacl::MAIN_MEMORY, std::size_t(cols));
  vec_t vy(y, viennacl::MAIN_MEMORY, std::size_t(rows));
  viennacl::linalg::prod_impl(M, vx, alpha, vy, beta);
  return 0;
}

// diag[r] = A(r,r) (0 when absent): detail::row_info(A, vec, SPARSE_ROW_DIAGONAL)
int vclref_csr_diag(int rows, int cols, int nnz, const u32 *rp, const u32 *ci, const real_t *v, real_t *diag)
{
  csr_t A(const_cast<u32*>(rp), const_cast<u32*>(ci), const_cast<real_t*>(v), viennacl::MAIN_MEMORY, rows, cols, nnz);
  vec_t d(diag, viennacl::MAIN_MEMORY, std::size_t(rows));
  viennacl::linalg::detail::row_info(A, d, viennacl::linalg::detail::SPARSE_ROW_DIAGONAL);
  return 0;
}

real_t vclref_norm2(const real_t *x, int n)
{
  vec_t vx(const_cast<real_t*>(x), viennacl::MAIN_MEMORY, std::size_t(n));
  real_t r = viennacl::linalg::norm_2(vx);
  return r;
}

real_t vclref_inner_prod(const real_t *x, const real_t *y, int n)
{
  vec_t vx(const_cast<real_t*>(x), viennacl::MAIN_MEMORY, std::size_t(n));
  vec_t vy(const_cast<real_t*>(y), viennacl::MAIN_MEMORY, std::size_t(n));
  real_t r = viennacl::linalg::inner_prod(vx, vy);
  return r;
}

// solver: 0 CG, 1 BiCGStab, 2 GMRES;  precond: 0 none (pipelined path), 1 Jacobi, 2 identity functor (generic path)
// format: 0 CSR, 1 SELL-32 (rows % 32 != 0 required), 2 ELL, 3 HYB
// hist (optional): monitor estimates, hist_len receives the number of monitor calls.
int vclref_solve(int solver, int precond, int format,
                 int rows, int nnz, const u32 *rp, const u32 *ci, const real_t *v,
                 const real_t *b, real_t *x,
                 double tol, double abs_tol, int maxit, int krylov, int restart_every,
                 int *iters, double *err, double *hist, int hist_cap, int *hist_len, double *seconds)
{
  csr_t A(const_cast<u32*>(rp), const_cast<u32*>(ci), const_cast<real_t*>(v), viennacl::MAIN_MEMORY, rows, rows, nnz);
  vec_t vb(const_cast<real_t*>(b), viennacl::MAIN_MEMORY, std::size_t(rows));
  vec_t vx(std::size_t(rows), viennacl::context(viennacl::MAIN_MEMORY));
  history h = {hist, hist_cap, 0};
  history *hp = (hist_len != NULL) ? &h : NULL;
  int rc;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  if (format == 0)
    rc = run_solver(A, solver, precond, &A, vb, vx, tol, abs_tol, maxit, krylov, restart_every, iters, err, hp);
  else if (format == 2 || format == 3)
  {
    raw_csr_view view = {std::size_t(rows), std::size_t(rows), rp, ci, v};
    if (format == 2)
    {
      ell_t E(viennacl::context(viennacl::MAIN_MEMORY));
      viennacl::copy(view, E);
      t0 = std::chrono::steady_clock::now();
      rc = run_solver(E, solver, precond, &A, vb, vx, tol, abs_tol, maxit, krylov, restart_every, iters, err, hp);
    }
    else
    {
      hyb_t H(viennacl::context(viennacl::MAIN_MEMORY));
      viennacl::copy(view, H);
      t0 = std::chrono::steady_clock::now();
      rc = run_solver(H, solver, precond, &A, vb, vx, tol, abs_tol, maxit, krylov, restart_every, iters, err, hp);
    }
  }
  else
  {
    if (rows % 32 == 0) return 3;
    raw_csr_view view = {std::size_t(rows), std::size_t(rows), rp, ci, v};
    sell_t S((std::size_t(rows)), (std::size_t(rows)), 32);
    viennacl::copy(view, S);
    t0 = std::chrono::steady_clock::now();
    rc = run_solver(S, solver, precond, &A, vb, vx, tol, abs_tol, maxit, krylov, restart_every, iters, err, hp);
  }
  std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  if (hist_len) *hist_len = h.len;
  if (rc == 0) viennacl::fast_copy(vx.begin(), vx.end(), x);
  return rc;
}

#ifndef VCLREF_F32
// mixed_precision_cg.hpp:95-186 on the host backend (double system, float inner iterations)
int vclref_mixed_cg(int rows, int nnz, const u32 *rp, const u32 *ci, const double *v, const double *b, double *x,
                    double tol, int maxit, float inner_tol, int *iters, double *err)
{
  csr_t A(const_cast<u32*>(rp), const_cast<u32*>(ci), const_cast<double*>(v), viennacl::MAIN_MEMORY, rows, rows, nnz);
  vec_t vb(const_cast<double*>(b), viennacl::MAIN_MEMORY, std::size_t(rows));
  viennacl::linalg::mixed_precision_cg_tag tag(tol, maxit, inner_tol);
  vec_t vx = viennacl::linalg::solve(A, vb, tag);
  *iters = int(tag.iters()); *err = tag.error();
  viennacl::fast_copy(vx.begin(), vx.end(), x);
  return 0;
}
#endif

// Times `reps` plain y = A*x products the way examples/benchmarks/sparse.cpp:123-130 does (one warm-up first).
double vclref_time_csr_spmv(int rows, int cols, int nnz, const u32 *rp, const u32 *ci, const real_t *v,
                            const real_t *x, real_t *y, int reps)
{
  csr_t A(const_cast<u32*>(rp), const_cast<u32*>(ci), const_cast<real_t*>(v), viennacl::MAIN_MEMORY, rows, cols, nnz);
  vec_t vx(const_cast<real_t*>(x), viennacl::MAIN_MEMORY, std::size_t(cols));
  vec_t vy(y, viennacl::MAIN_MEMORY, std::size_t(rows));
  vy = viennacl::linalg::prod(A, vx);
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i)
    vy = viennacl::linalg::prod(A, vx);
  std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
