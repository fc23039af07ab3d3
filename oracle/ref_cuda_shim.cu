// oracle/ref_cuda_shim.cu -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" shim over the UNMODIFIED reference headers under /root/reference, compiled with the reference's OWN CUDA
// backend (-DVIENNACL_WITH_CUDA, nvcc -arch=sm_100: "legacy kernels on B200").  The shared object lives in
// oracle/_ref/libvcl_ref_cuda.so (git-ignored, travels to the GPU box).  It is the SECOND baseline of bench.py
// (`legacy_cuda_baseline`): the reference's kernels
//     viennacl/linalg/cuda/sparse_matrix_operations.hpp:137-249, 262-396   CSR SpMV (K1/K2)
//     viennacl/linalg/cuda/sparse_matrix_operations.hpp:2196-2289         SELL SpMV (K4)
//     viennacl/linalg/cuda/iterative_operations.hpp                        fused pipelined CG / BiCGStab / GMRES kernels
//     viennacl/linalg/cg.hpp:128-187, bicgstab.hpp:97-215, gmres.hpp:181-367  drivers (one blocking D2H per iteration)
// run on the same GPU, on the same matrices, next to the repo's arm.  Results are returned so that the caller can check them
// against the oracle before trusting the timing (SURVEY 8c warns about the warp-synchronous K1).
// Only tests/ and bench.py's baseline legs may load it.
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
#include <cuda_runtime.h>

#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/sliced_ell_matrix.hpp"
#include "viennacl/linalg/prod.hpp"
#include "viennacl/linalg/cg.hpp"
#include "viennacl/linalg/bicgstab.hpp"
#include "viennacl/linalg/gmres.hpp"
#include "viennacl/linalg/jacobi_precond.hpp"

typedef unsigned int u32;
typedef viennacl::compressed_matrix<double> csr_t;
typedef viennacl::sliced_ell_matrix<double> sell_t;
typedef viennacl::vector<double> vec_t;

namespace {

// read-only iterator view over raw CSR arrays for the reference's generic copy(CPUMatrixT, sliced_ell_matrix)
struct raw_csr_view
{
  typedef std::size_t size_type; typedef double value_type;
  std::size_t rows_, cols_; const u32 *rp_, *ci_; const double *v_;
  struct const_iterator2
  {
    const raw_csr_view *m; std::size_t row, k;
    std::size_t index1() const { return row; }
    std::size_t index2() const { return m->ci_[k]; }
    double operator*() const { return m->v_[k]; }
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

double now_s()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

extern "C" {

// 1 when a CUDA device is usable by the reference backend
int vclrefcuda_available(void)
{
  int n = 0;
  return (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) ? 1 : 0;
}

// y = A x through the reference's CUDA backend; format 0: compressed_matrix, 1: sliced_ell_matrix (C = 32).
// One warm-up product, then `reps` products between two backend::finish() calls (examples/benchmarks/sparse.cpp:123-130 protocol).
// y_out (host, rows entries) receives the result; *seconds the time of the `reps` products.
int vclrefcuda_spmv(int format, int rows, int cols, int nnz, const u32 *rp, const u32 *ci, const double *v,
                    const double *x, double *y_out, int reps, double *seconds)
{
  try
  {
    viennacl::context ctx(viennacl::CUDA_MEMORY);
    vec_t vx(std::size_t(cols), ctx), vy(std::size_t(rows), ctx);
    viennacl::fast_copy(x, x + cols, vx.begin());
    double t0 = 0, t1 = 0;
    if (format == 0)
    {
      csr_t A(ctx);
      A.set(rp, ci, v, std::size_t(rows), std::size_t(cols), std::size_t(nnz));
      vy = viennacl::linalg::prod(A, vx);
      viennacl::backend::finish();
      t0 = now_s();
      for (int i = 0; i < reps; ++i) vy = viennacl::linalg::prod(A, vx);
      viennacl::backend::finish();
      t1 = now_s();
    }
    else
    {
      raw_csr_view view = {std::size_t(rows), std::size_t(cols), rp, ci, v};
      sell_t S((std::size_t(rows)), (std::size_t(cols)), 32);          // default memory domain = CUDA in this build (backend/mem_handle.hpp:51-60)
      viennacl::copy(view, S);
      vy = viennacl::linalg::prod(S, vx);
      viennacl::backend::finish();
      t0 = now_s();
      for (int i = 0; i < reps; ++i) vy = viennacl::linalg::prod(S, vx);
      viennacl::backend::finish();
      t1 = now_s();
    }
    if (seconds) *seconds = t1 - t0;
    viennacl::fast_copy(vy.begin(), vy.end(), y_out);
    return 0;
  }
  catch (std::exception const & e) { std::fprintf(stderr, "vclrefcuda_spmv: %s\n", e.what()); return 1; }
}

// solve(A, b, tag) with the reference's CUDA backend (compressed_matrix); solver 0 CG, 1 BiCGStab, 2 GMRES; precond 0 none
// (the pipelined paths), 1 Jacobi (the generic paths).  Timed like examples/benchmarks/solver.cpp:106-120 (finish, timer, solve, finish).
int vclrefcuda_solve(int solver, int precond, int rows, int nnz, const u32 *rp, const u32 *ci, const double *v,
                     const double *b, double *x_out, double tol, int maxit, int krylov,
                     int *iters, double *err, double *seconds)
{
  try
  {
    viennacl::context ctx(viennacl::CUDA_MEMORY);
    csr_t A(ctx);
    A.set(rp, ci, v, std::size_t(rows), std::size_t(rows), std::size_t(nnz));
    vec_t vb(std::size_t(rows), ctx), vx(std::size_t(rows), ctx);
    viennacl::fast_copy(b, b + rows, vb.begin());
    viennacl::backend::finish();
    const double t0 = now_s();
    if (solver == 0)
    {
      viennacl::linalg::cg_tag tag(tol, maxit);
      if (precond == 1) { viennacl::linalg::jacobi_precond<csr_t> jac(A, viennacl::linalg::jacobi_tag()); vx = viennacl::linalg::solve(A, vb, tag, jac); }
      else vx = viennacl::linalg::solve(A, vb, tag);
      *iters = int(tag.iters()); *err = tag.error();
    }
    else if (solver == 1)
    {
      viennacl::linalg::bicgstab_tag tag(tol, maxit);
      if (precond == 1) { viennacl::linalg::jacobi_precond<csr_t> jac(A, viennacl::linalg::jacobi_tag()); vx = viennacl::linalg::solve(A, vb, tag, jac); }
      else vx = viennacl::linalg::solve(A, vb, tag);
      *iters = int(tag.iters()); *err = tag.error();
    }
    else
    {
      viennacl::linalg::gmres_tag tag(tol, maxit, krylov);
      if (precond == 1) { viennacl::linalg::jacobi_precond<csr_t> jac(A, viennacl::linalg::jacobi_tag()); vx = viennacl::linalg::solve(A, vb, tag, jac); }
      else vx = viennacl::linalg::solve(A, vb, tag);
      *iters = int(tag.iters()); *err = tag.error();
    }
    viennacl::backend::finish();
    if (seconds) *seconds = now_s() - t0;
    viennacl::fast_copy(vx.begin(), vx.end(), x_out);
    return 0;
  }
  catch (std::exception const & e) { std::fprintf(stderr, "vclrefcuda_solve: %s\n", e.what()); return 1; }
}

} // extern "C"
