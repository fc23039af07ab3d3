// bench_sparse_solver.cpp -- the reference's benchmark programs (examples/benchmarks/sparse.cpp:57-230 and
// solver.cpp:101-125, 160-640) for the formats and solvers of the hot path, against the B200 facade, without uBLAS and
// without the missing mat65k.mtx fixture: FDM matrices generated on the device.  Timing protocol is the reference's:
// warm-up call, backend::finish(), wall-clock timer around BENCHMARK_RUNS calls, finish().  Usage: bench_sparse_solver [n3d] [n2d]
#include <cstdlib>
#include <iostream>
#include <vector>

#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/sliced_ell_matrix.hpp"
#include "viennacl/linalg/prod.hpp"
#include "viennacl/linalg/norm_2.hpp"
#include "viennacl/linalg/jacobi_precond.hpp"
#include "viennacl/linalg/cg.hpp"
#include "viennacl/linalg/bicgstab.hpp"
#include "viennacl/linalg/gmres.hpp"
#include "viennacl/tools/matrix_generation.hpp"
#include "viennacl/tools/timer.hpp"

#define BENCHMARK_RUNS 10
typedef double ScalarType;
typedef viennacl::vector<ScalarType> VectorT;

static void printOps(double num_ops, double exec_time) { std::cout << "GFLOPs: " << num_ops / (1000000 * exec_time * 1000) << std::endl; }

template<typename MatrixT>
static void bench_prod(const char *name, MatrixT const & A, double nnz, double bytes, VectorT & y, VectorT const & x)
{
  viennacl::tools::timer timer;
  std::cout << "------- Matrix-Vector product with " << name << " ----------" << std::endl;
  y = viennacl::linalg::prod(A, x);                 // startup calculation
  viennacl::backend::finish();
  timer.start();
  for (int runs = 0; runs < BENCHMARK_RUNS; ++runs) y = viennacl::linalg::prod(A, x);
  viennacl::backend::finish();
  double exec_time = timer.get();
  std::cout << "GPU time align1: " << exec_time << std::endl;
  std::cout << "GPU align1 "; printOps(2.0 * nnz, exec_time / BENCHMARK_RUNS);
  std::cout << "effective GB/s: " << bytes * BENCHMARK_RUNS / exec_time * 1e-9 << std::endl;
  std::cout << y[0] << std::endl;
}

template<typename MatrixT, typename SolverTag, typename PrecondT>
static void run_solver(const char *name, MatrixT const & matrix, VectorT const & rhs, SolverTag const & solver, PrecondT const & precond)
{
  viennacl::tools::timer timer;
  std::cout << "------- " << name << " ----------" << std::endl;
  VectorT result(rhs.size());
  viennacl::backend::finish();
  timer.start();
  result = viennacl::linalg::solve(matrix, rhs, solver, precond);
  viennacl::backend::finish();
  double exec_time = timer.get();
  std::cout << "Exec. time: " << exec_time << std::endl;
  VectorT residual = rhs - viennacl::linalg::prod(matrix, result);
  std::cout << "Relative residual: " << ScalarType(viennacl::linalg::norm_2(residual)) / ScalarType(viennacl::linalg::norm_2(rhs)) << std::endl;
  std::cout << "Estimated rel. residual: " << solver.error() << std::endl;
  std::cout << "Iterations: " << solver.iters() << "  (" << solver.iters() / exec_time << " iterations/s)" << std::endl;
}

int main(int argc, char **argv)
{
  std::size_t n3 = argc > 1 ? std::size_t(std::atoi(argv[1])) : 128, n2 = argc > 2 ? std::size_t(std::atoi(argv[2])) : 512;
  viennacl::compressed_matrix<ScalarType> A3, A2, C3;
  viennacl::tools::generate_fdm_stencil(A3, n3, n3, n3);
  viennacl::tools::generate_fdm_stencil(C3, n3, n3, n3, 0.5, 0.25, 0.125);
  viennacl::tools::generate_fdm_laplace(A2, n2, n2);
  viennacl::sliced_ell_matrix<ScalarType> S3;
  viennacl::copy(A3, S3);
  std::cout << "3-D 7-point " << n3 << "^3: " << A3.size1() << " rows, " << A3.nnz() << " nonzeros; 2-D 5-point " << n2 << "^2: " << A2.nnz() << " nonzeros" << std::endl;

  VectorT x = viennacl::scalar_vector<ScalarType>(A3.size1(), 1.0), y(A3.size1());
  double N = double(A3.size1()), nnz = double(A3.nnz());
  bench_prod("compressed_matrix", A3, nnz, 12.0 * nnz + 20.0 * N, y, x);
  bench_prod("sliced_ell_matrix", S3, nnz, 12.0 * nnz + 16.0 * N, y, x);

  VectorT b2 = viennacl::scalar_vector<ScalarType>(A2.size1(), 1.0), b3 = x;
  viennacl::linalg::no_precond none;
  viennacl::linalg::jacobi_precond< viennacl::compressed_matrix<ScalarType> > jacobi(C3, viennacl::linalg::jacobi_tag());
  run_solver("CG solver, 2-D Laplacian, compressed_matrix", A2, b2, viennacl::linalg::cg_tag(1e-8, 5000), none);
  run_solver("CG solver, 3-D Laplacian, sliced_ell_matrix", S3, b3, viennacl::linalg::cg_tag(1e-8, 5000), none);
  run_solver("BiCGStab solver, 3-D convection-diffusion, no preconditioner", C3, b3, viennacl::linalg::bicgstab_tag(1e-8, 2000), none);
  run_solver("BiCGStab solver, 3-D convection-diffusion, Jacobi preconditioner", C3, b3, viennacl::linalg::bicgstab_tag(1e-8, 2000), jacobi);
  run_solver("GMRES(30) solver, 3-D convection-diffusion", C3, b3, viennacl::linalg::gmres_tag(1e-8, 900, 30), none);
  return EXIT_SUCCESS;
}
