// iterative.cpp -- the user code of the reference's solver tutorials (examples/tutorial/iterative.cpp:190-290 and
// iterative-custom.cpp:55-200) against the B200 facade: solve(A, b, tag[, precond]) for CG / BiCGStab / GMRES on
// compressed_matrix and sliced_ell_matrix, the *_solver functors with set_initial_guess() and set_monitor().
// The tutorials' fixture mat65k.mtx is not shipped; the systems are the FDM matrices of SURVEY 8(d) at small sizes.
// Each result is verified by its TRUE relative residual ||b - A x|| / ||b|| computed with prod() + norm_2().
#include <cstdlib>
#include <cmath>
#include <iostream>
#include <vector>

#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/sliced_ell_matrix.hpp"
#include "viennacl/ell_matrix.hpp"
#include "viennacl/hyb_matrix.hpp"
#include "viennacl/coordinate_matrix.hpp"
#include "viennacl/linalg/prod.hpp"
#include "viennacl/linalg/norm_2.hpp"
#include "viennacl/linalg/jacobi_precond.hpp"
#include "viennacl/linalg/cg.hpp"
#include "viennacl/linalg/bicgstab.hpp"
#include "viennacl/linalg/gmres.hpp"
#include "viennacl/linalg/mixed_precision_cg.hpp"
#include "viennacl/tools/matrix_generation.hpp"

// built twice: iterative (double) and iterative_float (-DNUMERIC_T=float; the reference's tutorials run with either)
#ifndef NUMERIC_T
#define NUMERIC_T double
#endif
typedef NUMERIC_T ScalarType;
static const bool SINGLE = sizeof(ScalarType) == 4;
static const double TOL = SINGLE ? 1e-5 : 1e-8;      // solver tolerances a float solve can reach
static const double TOL10 = SINGLE ? 1e-5 : 1e-10;
// bounds on the true relative residual.  float: x itself is only known to 6e-8 relative, so ||b - A x|| / ||b|| cannot fall
// below ~ eps * ||A|| ||x|| / ||b|| (2..4e-4 on these systems; the float reference stalls at the same level)
static const double RES7 = SINGLE ? 1e-3 : 1e-7;
static const double RES6 = SINGLE ? 1e-3 : 1e-6;
static const double MON = SINGLE ? 1e-3 : 1e-4;      // the custom monitor's own stopping threshold
typedef viennacl::vector<ScalarType> VectorT;
typedef viennacl::compressed_matrix<ScalarType> MatrixT;

template<typename MatT>
static ScalarType true_residual(MatT const & A, VectorT const & x, VectorT const & b)
{
  VectorT r = viennacl::linalg::prod(A, x);
  r = b - r;
  return ScalarType(viennacl::linalg::norm_2(r)) / ScalarType(viennacl::linalg::norm_2(b));
}

static int failures = 0;
static void expect(bool ok, const char *what)
{
  std::cout << (ok ? "  ok  " : "# FAILED: ") << what << std::endl;
  if (!ok) ++failures;
}

// iterative-custom.cpp:55-100
template<typename MatT>
struct monitor_user_data
{
  monitor_user_data(MatT const & A, VectorT const & b, VectorT const & guess) : A_ptr(&A), b_ptr(&b), guess_ptr(&guess), calls(0), last_true(0), last_est(0) {}
  MatT const *A_ptr;
  VectorT const *b_ptr;
  VectorT const *guess_ptr;
  int calls;
  ScalarType last_true, last_est;
};

template<typename MatT>
bool my_custom_monitor(VectorT const & current_approx, ScalarType residual_estimate, void *user_data)
{
  monitor_user_data<MatT> *data = reinterpret_cast<monitor_user_data<MatT>*>(user_data);
  VectorT x = current_approx + *data->guess_ptr;          // the solver works on the shifted system
  data->last_true = true_residual(*data->A_ptr, x, *data->b_ptr);
  data->last_est = residual_estimate;
  ++data->calls;
  return data->last_true < MON;                           // custom termination criterion
}

int main()
{
  MatrixT A;                                               // 2-D Laplacian, SPD
  viennacl::tools::generate_fdm_laplace(A, 96, 80);
  MatrixT C;                                               // 3-D upwind convection-diffusion, nonsymmetric
  viennacl::tools::generate_fdm_stencil(C, 24, 20, 18, 0.5, 0.25, 0.125);
  viennacl::sliced_ell_matrix<ScalarType> A_sell, C_sell;
  viennacl::copy(A, A_sell);
  viennacl::copy(C, C_sell);

  VectorT b = viennacl::scalar_vector<ScalarType>(A.size1(), 1.0);
  VectorT c = viennacl::scalar_vector<ScalarType>(C.size1(), 1.0);

  std::cout << "----- CG Method -----" << std::endl;
  {
    VectorT x;
    viennacl::linalg::cg_tag tag(TOL, 1000);
    x = viennacl::linalg::solve(A, b, tag);
    std::cout << "  CSR : " << tag.iters() << " iterations, estimate " << tag.error() << ", true " << true_residual(A, x, b) << std::endl;
    expect(tag.iters() > 10 && tag.iters() < 1000 && tag.error() < TOL && true_residual(A, x, b) < RES7, "solve(compressed_matrix, b, cg_tag)");
    viennacl::linalg::cg_tag tag2(TOL, 1000);
    VectorT x2 = viennacl::linalg::solve(A_sell, b, tag2);
    expect(std::abs(int(tag2.iters()) - int(tag.iters())) <= 2 && true_residual(A, x2, b) < RES7, "solve(sliced_ell_matrix, b, cg_tag)");
    viennacl::linalg::cg_tag few(TOL, 20);
    x = viennacl::linalg::solve(A, b, few);
    expect(few.iters() == 20 && few.error() > TOL, "cg_tag(TOL, 20) stops at max_iterations and reports the estimate");
    x = viennacl::linalg::solve(A, b, viennacl::linalg::cg_tag(), viennacl::linalg::no_precond());
    expect(x.size() == b.size(), "solve(A, b, cg_tag(), no_precond())");
  }

  std::cout << "----- mixed-precision CG (mixed_precision_cg.hpp) -----" << std::endl;
  if (!SINGLE)
  {
    viennacl::compressed_matrix<double> Ad;
    viennacl::tools::generate_fdm_laplace(Ad, 96, 80);
    viennacl::vector<double> bd = viennacl::scalar_vector<double>(Ad.size1(), 1.0);
    viennacl::linalg::mixed_precision_cg_tag mtag(1e-8, 2000, 1e-2f);
    viennacl::vector<double> xm = viennacl::linalg::solve(Ad, bd, mtag);
    viennacl::vector<double> rm = viennacl::linalg::prod(Ad, xm);
    rm = bd - rm;
    const double true_res = viennacl::linalg::norm_2(rm) / viennacl::linalg::norm_2(bd);
    std::cout << "  mixed CG : " << mtag.iters() << " iterations, error " << mtag.error() << ", true " << true_res << std::endl;
    expect(mtag.iters() > 10 && mtag.iters() < 2000 && mtag.error() < 1e-8 && true_res < 1e-8 && std::fabs(true_res - mtag.error()) < 1e-9,
           "solve(compressed_matrix<double>, b, mixed_precision_cg_tag): double-accurate result from float inner iterations");
  }

  std::cout << "----- BiCGStab Method -----" << std::endl;
  {
    VectorT x;
    viennacl::linalg::bicgstab_tag tag(TOL, 1000);
    x = viennacl::linalg::solve(C, c, tag);
    std::cout << "  CSR : " << tag.iters() << " iterations, estimate " << tag.error() << ", true " << true_residual(C, x, c) << std::endl;
    expect(tag.iters() > 3 && tag.iters() < 1000 && true_residual(C, x, c) < RES6, "solve(compressed_matrix, b, bicgstab_tag)");
    viennacl::linalg::bicgstab_tag tag2(TOL, 1000);
    VectorT x2 = viennacl::linalg::solve(C_sell, c, tag2);
    expect(tag2.iters() < 1000 && true_residual(C, x2, c) < RES6, "solve(sliced_ell_matrix, b, bicgstab_tag)");
    viennacl::linalg::jacobi_precond<MatrixT> vcl_jacobi(C, viennacl::linalg::jacobi_tag());
    viennacl::linalg::bicgstab_tag tag3(TOL, 1000);
    x = viennacl::linalg::solve(C, c, tag3, vcl_jacobi);
    std::cout << "  CSR + Jacobi : " << tag3.iters() << " iterations, estimate " << tag3.error() << ", true " << true_residual(C, x, c) << std::endl;
    expect(tag3.iters() > 3 && tag3.iters() < 1000 && true_residual(C, x, c) < RES6, "solve(compressed_matrix, b, bicgstab_tag, jacobi_precond)");
    VectorT y = c;
    vcl_jacobi.apply(y);                                   // stand-alone apply: y = c ./ diag(C)
    expect(std::fabs(ScalarType(y[5]) - 1.0 / (6.0 + 0.5 + 0.25 + 0.125)) < (SINGLE ? 1e-7 : 1e-15), "jacobi_precond::apply");
  }

  std::cout << "----- GMRES Method -----" << std::endl;
  {
    VectorT x;
    viennacl::linalg::gmres_tag tag(TOL, 600, 30);
    x = viennacl::linalg::solve(C, c, tag);
    std::cout << "  CSR : " << tag.iters() << " iterations, estimate " << tag.error() << ", true " << true_residual(C, x, c) << std::endl;
    expect(tag.iters() > 3 && tag.iters() < 600 && true_residual(C, x, c) < RES7, "solve(compressed_matrix, b, gmres_tag)");
    viennacl::linalg::gmres_tag tag2(TOL, 600, 30);
    VectorT x2 = viennacl::linalg::solve(C_sell, c, tag2);
    expect((SINGLE ? std::abs(int(tag2.iters()) - int(tag.iters())) <= 60 : tag2.iters() == tag.iters()) && true_residual(C, x2, c) < RES7, "solve(sliced_ell_matrix, b, gmres_tag)");
    expect(viennacl::linalg::gmres_tag(TOL, 90, 30).max_restarts() == 2 && viennacl::linalg::gmres_tag(TOL, 100, 30).max_restarts() == 3,
           "gmres_tag::max_restarts() (gmres.hpp:74-80)");
  }

  std::cout << "----- ell_matrix / hyb_matrix in the solvers (cg.hpp:204-254 overloads) -----" << std::endl;
  {
    viennacl::ell_matrix<ScalarType> A_ell, C_ell;
    viennacl::hyb_matrix<ScalarType> A_hyb, C_hyb;
    viennacl::copy(A, A_ell); viennacl::copy(C, C_ell);
    viennacl::copy(A, A_hyb);
    C_hyb.csr_threshold(0.2);                              // ELL width 6: the 7-entry interior rows spill into the CSR tail
    viennacl::copy(C, C_hyb);
    viennacl::linalg::cg_tag ref_tag(TOL, 1000), t1(TOL, 1000), t2(TOL, 1000);
    VectorT x0 = viennacl::linalg::solve(A, b, ref_tag);
    VectorT x1 = viennacl::linalg::solve(A_ell, b, t1);
    VectorT x2 = viennacl::linalg::solve(A_hyb, b, t2);
    expect(std::abs(int(t1.iters()) - int(ref_tag.iters())) <= 2 && std::abs(int(t2.iters()) - int(ref_tag.iters())) <= 2 &&
           true_residual(A, x1, b) < RES7 && true_residual(A, x2, b) < RES7, "solve(ell_matrix / hyb_matrix, b, cg_tag)");
    viennacl::linalg::bicgstab_tag t3(TOL, 1000), t4(TOL, 1000);
    VectorT x3 = viennacl::linalg::solve(C_ell, c, t3);
    VectorT x4 = viennacl::linalg::solve(C_hyb, c, t4);
    expect(C_hyb.csr_nnz() > 1 && true_residual(C, x3, c) < RES6 && true_residual(C, x4, c) < RES6, "solve(ell_matrix / hyb_matrix, b, bicgstab_tag)");
    viennacl::linalg::gmres_tag t5(TOL, 600, 30), t6(TOL, 600, 30);
    VectorT x5 = viennacl::linalg::solve(C_ell, c, t5);
    VectorT x6 = viennacl::linalg::solve(C_hyb, c, t6);
    expect((SINGLE ? std::abs(int(t5.iters()) - int(t6.iters())) <= 60 : t5.iters() == t6.iters()) && true_residual(C, x5, c) < RES7 && true_residual(C, x6, c) < RES7, "solve(ell_matrix / hyb_matrix, b, gmres_tag)");
    // coordinate_matrix (iterative.cpp:126-128, 160: the tutorial copies the system into a coordinate_matrix as well)
    std::vector< std::map<unsigned int, ScalarType> > stl_A;
    viennacl::copy(A, stl_A);
    viennacl::coordinate_matrix<ScalarType> A_coo;
    viennacl::copy(stl_A, A_coo);
    viennacl::linalg::cg_tag t7(TOL, 1000);
    VectorT x7 = viennacl::linalg::solve(A_coo, b, t7);
    expect(std::abs(int(t7.iters()) - int(ref_tag.iters())) <= 2 && true_residual(A, x7, b) < RES7 && A_coo.nnz() == A.nnz(),
           "solve(coordinate_matrix, b, cg_tag)");
  }

  std::cout << "----- solver objects: initial guess + monitor (iterative-custom.cpp) -----" << std::endl;
  {
    VectorT init_guess = viennacl::scalar_vector<ScalarType>(b.size(), 0.9);
    init_guess[0] = 0;
    monitor_user_data<MatrixT> data(A, b, init_guess);

    viennacl::linalg::cg_solver<VectorT> my_cg_solver(viennacl::linalg::cg_tag(TOL10, 2000));
    my_cg_solver.set_monitor(my_custom_monitor<MatrixT>, &data);
    my_cg_solver.set_initial_guess(init_guess);
    VectorT x = my_cg_solver(A, b);
    std::cout << "  CG : monitor called " << data.calls << " times, stopped at true residual " << data.last_true << " (estimate " << data.last_est << ")" << std::endl;
    expect(data.calls > 5 && data.last_true < MON && true_residual(A, x, b) < MON && true_residual(A, x, b) > 1e-9, "cg_solver with monitor stops early");

    monitor_user_data<MatrixT> data2(C, c, c);
    VectorT guess2 = viennacl::scalar_vector<ScalarType>(c.size(), 0.1);
    data2.guess_ptr = &guess2;
    viennacl::linalg::bicgstab_solver<VectorT> my_bicgstab_solver(viennacl::linalg::bicgstab_tag(TOL10, 2000));
    my_bicgstab_solver.set_monitor(my_custom_monitor<MatrixT>, &data2);
    my_bicgstab_solver.set_initial_guess(guess2);
    VectorT xc = my_bicgstab_solver(C, c);
    expect(data2.calls > 1 && true_residual(C, xc, c) < 1e-3, "bicgstab_solver with monitor + initial guess");

    monitor_user_data<MatrixT> data3(C, c, guess2);
    viennacl::linalg::gmres_solver<VectorT> my_gmres_solver(viennacl::linalg::gmres_tag(TOL10, 600, 20));
    my_gmres_solver.set_monitor(my_custom_monitor<MatrixT>, &data3);
    my_gmres_solver.set_initial_guess(guess2);
    xc = my_gmres_solver(C, c);
    std::cout << "  GMRES : monitor called " << data3.calls << " times (once per restart)" << std::endl;
    expect(data3.calls >= 1 && true_residual(C, xc, c) < MON, "gmres_solver with monitor + initial guess");

    viennacl::linalg::cg_solver<VectorT> plain(viennacl::linalg::cg_tag(TOL, 1000));
    plain.set_initial_guess(init_guess);
    x = plain(A, b);
    expect(true_residual(A, x, b) < RES6, "cg_solver with initial guess, no monitor");
  }

  if (failures) { std::cout << failures << " check(s) FAILED" << std::endl; return EXIT_FAILURE; }
  std::cout << "!!!! TUTORIAL COMPLETED SUCCESSFULLY !!!!" << std::endl;
  return EXIT_SUCCESS;
}
