// matrix_free.cpp -- examples/tutorial/matrix-free.cpp (user operators with apply()/size1() in all three solvers) and the
// preconditioned calls of examples/tutorial/iterative.cpp:215-290 (solve(A, b, tag, precond) with Jacobi and with a
// user-defined preconditioner) against the B200 facade.  These go through the GENERIC solver paths of the facade
// (cg.hpp:257-322, bicgstab.hpp:398-489, gmres.hpp:449-631 in the reference), built on prod / inner_prod / norm_2.
#include <cstdlib>
#include <cmath>
#include <iostream>
#include <map>
#include <vector>

#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/linalg/prod.hpp"
#include "viennacl/linalg/norm_2.hpp"
#include "viennacl/linalg/jacobi_precond.hpp"
#include "viennacl/linalg/row_scaling.hpp"
#include "viennacl/linalg/cg.hpp"
#include "viennacl/linalg/bicgstab.hpp"
#include "viennacl/linalg/gmres.hpp"
#include "viennacl/tools/matrix_generation.hpp"

typedef double ScalarType;
typedef viennacl::vector<ScalarType> VectorT;
typedef viennacl::compressed_matrix<ScalarType> MatrixT;

// A user-defined operator: y = (L + sigma I) x, never assembled (matrix-free.cpp:38-75)
template<typename NumericT>
class MyOperator
{
public:
  MyOperator(viennacl::compressed_matrix<NumericT> const & L, NumericT sigma) : L_(&L), sigma_(sigma) {}
  void apply(viennacl::vector_base<NumericT> const & x, viennacl::vector_base<NumericT> & y) const
  {
    y = viennacl::linalg::prod(*L_, x);
    y += sigma_ * x;
  }
  std::size_t size1() const { return L_->size1(); }
private:
  viennacl::compressed_matrix<NumericT> const *L_;
  NumericT sigma_;
};

// A user-defined preconditioner: anything with apply(v) (iterative.cpp uses ILU objects this way)
struct MyDiagonalScaling
{
  explicit MyDiagonalScaling(VectorT const & inv_diag) : inv_diag_(&inv_diag) {}
  void apply(VectorT & v) const { VectorT t = v; v = viennacl::linalg::element_div(t, *inv_diag_); }
  VectorT const *inv_diag_;
};

template<typename OpT>
static ScalarType true_residual(OpT const & A, VectorT const & x, VectorT const & b)
{
  VectorT r = b - viennacl::linalg::prod(A, x);
  return ScalarType(viennacl::linalg::norm_2(r)) / ScalarType(viennacl::linalg::norm_2(b));
}

// Iteration counts of the UNMODIFIED reference (host backend, generic solver paths) on exactly these systems:
// tests/golden/generic_solver_counts.json, produced by oracle/ref_generic_counts.cpp (make -C oracle generic_counts).
static const int REF_OP_CG = 54, REF_OP_BICGSTAB = 40, REF_OP_GMRES30 = 56, REF_VARDIAG_CG = 154, REF_VARDIAG_CG_JACOBI = 7,
                 REF_CD3D_BICGSTAB_JACOBI = 49, REF_CD3D_GMRES20_JACOBI = 114, REF_CD3D_GMRES20_HOUSEHOLDER = 122,
                 REF_VARDIAG_CG_ROWSCALING2 = 7, REF_CD3D_BICGSTAB_ROWSCALING1 = 49, REF_CD3D_GMRES20_ROWSCALING0 = 114;
static bool within(unsigned int got, int ref, int tol) { return std::abs(int(got) - ref) <= tol; }

static int failures = 0;
static void expect(bool ok, const char *what)
{
  std::cout << (ok ? "  ok  " : "# FAILED: ") << what << std::endl;
  if (!ok) ++failures;
}

int main()
{
  // ---------------------------------------------------------------- matrix-free operators
  MatrixT L;
  viennacl::tools::generate_fdm_laplace(L, 48, 40);
  MyOperator<ScalarType> op(L, 0.25);
  VectorT rhs = viennacl::scalar_vector<ScalarType>(op.size1(), ScalarType(-1));

  std::cout << "----- user operator in CG / BiCGStab / GMRES -----" << std::endl;
  {
    viennacl::linalg::cg_tag tag(1e-9, 1000);
    VectorT x = viennacl::linalg::solve(op, rhs, tag);
    std::cout << "  CG: " << tag.iters() << " iterations, estimate " << tag.error() << ", true " << true_residual(op, x, rhs) << std::endl;
    expect(within(tag.iters(), REF_OP_CG, 2) && true_residual(op, x, rhs) < 1e-8, "solve(MyOperator, b, cg_tag): iterations within +-2 of the reference");
    viennacl::linalg::bicgstab_tag tag2(1e-9, 1000);
    VectorT x2 = viennacl::linalg::solve(op, rhs, tag2);
    expect(within(tag2.iters(), REF_OP_BICGSTAB, 4) && true_residual(op, x2, rhs) < 1e-7, "solve(MyOperator, b, bicgstab_tag): iterations within +-4 of the reference");
    viennacl::linalg::gmres_tag tag3(1e-9, 600, 30);
    VectorT x3 = viennacl::linalg::solve(op, rhs, tag3);
    std::cout << "  GMRES(30): " << tag3.iters() << " iterations, estimate " << tag3.error() << ", true " << true_residual(op, x3, rhs) << std::endl;
    expect(within(tag3.iters(), REF_OP_GMRES30, 2) && true_residual(op, x3, rhs) < 1e-8 && std::fabs(tag3.error() - true_residual(op, x3, rhs)) < 1e-9,
           "solve(MyOperator, b, gmres_tag): iterations within +-2 of the reference (Householder), estimate equals the true residual");
    VectorT diff = x - x3;
    expect(ScalarType(viennacl::linalg::norm_2(diff)) < 1e-6 * ScalarType(viennacl::linalg::norm_2(x)), "CG and GMRES agree on the solution");
  }

  // ---------------------------------------------------------------- preconditioned solves on assembled matrices
  // SPD matrix with a strongly varying diagonal: 2-D Laplacian + diag(1 .. 1000)
  const std::size_t nx = 40, ny = 36, n = nx * ny;
  std::vector< std::map<unsigned int, ScalarType> > stl_A(n);
  std::vector<ScalarType> inv_scaling(n);
  for (std::size_t j = 0; j < ny; ++j)
    for (std::size_t i = 0; i < nx; ++i)
    {
      const unsigned int r = static_cast<unsigned int>(i + nx * j);
      const ScalarType d = 4.0 + 1.0 + 999.0 * ScalarType((r * 7919u) % 1000u) / 999.0;
      stl_A[r][r] = d;
      inv_scaling[r] = d;
      if (i > 0) stl_A[r][r - 1] = -1.0;
      if (i + 1 < nx) stl_A[r][r + 1] = -1.0;
      if (j > 0) stl_A[r][r - static_cast<unsigned int>(nx)] = -1.0;
      if (j + 1 < ny) stl_A[r][r + static_cast<unsigned int>(nx)] = -1.0;
    }
  MatrixT A;
  viennacl::copy(stl_A, A);
  VectorT b = viennacl::scalar_vector<ScalarType>(n, 1.0);
  viennacl::linalg::jacobi_precond<MatrixT> vcl_jacobi(A, viennacl::linalg::jacobi_tag());

  std::cout << "----- CG with preconditioners -----" << std::endl;
  {
    viennacl::linalg::cg_tag plain(1e-10, 2000), jac(1e-10, 2000), user(1e-10, 2000);
    VectorT x0 = viennacl::linalg::solve(A, b, plain);
    VectorT x1 = viennacl::linalg::solve(A, b, jac, vcl_jacobi);
    VectorT diag(n);
    viennacl::copy(inv_scaling, diag);
    VectorT x2 = viennacl::linalg::solve(A, b, user, MyDiagonalScaling(diag));
    std::cout << "  iterations: none " << plain.iters() << ", Jacobi " << jac.iters() << ", user diagonal scaling " << user.iters() << std::endl;
    expect(true_residual(A, x1, b) < 1e-8 && within(jac.iters(), REF_VARDIAG_CG_JACOBI, 2) && within(plain.iters(), REF_VARDIAG_CG, 2),
           "solve(A, b, cg_tag[, jacobi_precond]): iterations within +-2 of the reference");
    expect(within(user.iters(), REF_VARDIAG_CG_JACOBI, 0) && std::abs(int(user.iters()) - int(jac.iters())) <= 1 && true_residual(A, x2, b) < 1e-8,
           "generic PCG with a user preconditioner equal to Jacobi: the reference's count; fused Jacobi-PCG within 1 of it");
    VectorT diff = x0 - x1;
    expect(ScalarType(viennacl::linalg::norm_2(diff)) < 1e-7 * ScalarType(viennacl::linalg::norm_2(x0)), "same solution with and without preconditioner");
    viennacl::linalg::row_scaling<MatrixT> rs2(A, viennacl::linalg::row_scaling_tag(2));        // fused path, scaling = row 2-norms
    viennacl::linalg::cg_tag rs_tag(1e-10, 2000);
    VectorT x4 = viennacl::linalg::solve(A, b, rs_tag, rs2);
    expect(true_residual(A, x4, b) < 1e-8 && within(rs_tag.iters(), REF_VARDIAG_CG_ROWSCALING2, 2),
           "solve(A, b, cg_tag, row_scaling(2)): iterations within +-2 of the reference");
    viennacl::linalg::cg_tag few(1e-6, 20);                 // iterative.cpp:222 -- cg_tag(1e-6, 20) with a preconditioner
    VectorT x3 = viennacl::linalg::solve(A, b, few, vcl_jacobi);
    expect(few.iters() <= 20 && x3.size() == n, "cg_tag(1e-6, 20) with Jacobi respects the iteration budget");
  }

  std::cout << "----- BiCGStab / GMRES with preconditioners -----" << std::endl;
  {
    MatrixT C;
    viennacl::tools::generate_fdm_stencil(C, 20, 18, 16, 0.5, 0.25, 0.125);
    VectorT c = viennacl::scalar_vector<ScalarType>(C.size1(), 1.0);
    viennacl::linalg::jacobi_precond<MatrixT> jacobi_C(C, viennacl::linalg::jacobi_tag());
    VectorT diagC = jacobi_C.diagonal();

    viennacl::linalg::bicgstab_tag fused(1e-9, 1000), generic(1e-9, 1000);
    VectorT x1 = viennacl::linalg::solve(C, c, fused, jacobi_C);                    // fused Jacobi path of the backend
    VectorT x2 = viennacl::linalg::solve(C, c, generic, MyDiagonalScaling(diagC));  // generic path, same mathematics
    std::cout << "  BiCGStab: fused Jacobi " << fused.iters() << " iterations, generic " << generic.iters() << std::endl;
    expect(true_residual(C, x1, c) < 1e-7 && true_residual(C, x2, c) < 1e-7 && within(fused.iters(), REF_CD3D_BICGSTAB_JACOBI, 4) &&
           within(generic.iters(), REF_CD3D_BICGSTAB_JACOBI, 4), "fused and generic left-preconditioned BiCGStab: iterations within +-4 of the reference");

    viennacl::linalg::gmres_tag g0(1e-9, 600, 20), g1(1e-9, 600, 20);
    VectorT y0 = viennacl::linalg::solve(C, c, g0);
    VectorT y1 = viennacl::linalg::solve(C, c, g1, jacobi_C);
    std::cout << "  GMRES(20): none " << g0.iters() << " iterations, Jacobi " << g1.iters() << std::endl;
    expect(true_residual(C, y1, c) < 1e-7 && within(g1.iters(), REF_CD3D_GMRES20_JACOBI, 2), "solve(A, b, gmres_tag, jacobi_precond): iterations within +-2 of the reference");
    viennacl::linalg::row_scaling<MatrixT> rs1(C, viennacl::linalg::row_scaling_tag(1)), rs0(C, viennacl::linalg::row_scaling_tag(0));
    viennacl::linalg::bicgstab_tag b_rs(1e-9, 1000);
    viennacl::linalg::gmres_tag g_rs(1e-9, 600, 20);
    VectorT z1 = viennacl::linalg::solve(C, c, b_rs, rs1);
    VectorT z2 = viennacl::linalg::solve(C, c, g_rs, rs0);
    std::cout << "  row_scaling: BiCGStab(p=1) " << b_rs.iters() << ", GMRES(20)(p=0) " << g_rs.iters() << " iterations" << std::endl;
    expect(true_residual(C, z1, c) < 1e-7 && within(b_rs.iters(), REF_CD3D_BICGSTAB_ROWSCALING1, 4) &&
           true_residual(C, z2, c) < 1e-7 && within(g_rs.iters(), REF_CD3D_GMRES20_ROWSCALING0, 2),
           "solve(A, b, bicgstab_tag / gmres_tag, row_scaling): iterations within +-4 / +-2 of the reference");
    VectorT probe = viennacl::scalar_vector<ScalarType>(C.size1(), 1.0);
    rs0.apply(probe);                                         // interior row: sup-norm = diagonal 6 + 0.5 + 0.25 + 0.125
    expect(std::fabs(ScalarType(probe[C.size1() / 2 + 7]) - 1.0 / 6.875) < 1e-15, "row_scaling::apply divides by the row norm");
    expect(int(g0.iters()) == ((REF_CD3D_GMRES20_HOUSEHOLDER + 19) / 20) * 20,
           "pipelined GMRES(20) stops at the restart boundary after the reference's Householder count (gmres.hpp:234)");
  }

  if (failures) { std::cout << failures << " check(s) FAILED" << std::endl; return EXIT_FAILURE; }
  std::cout << "!!!! TUTORIAL COMPLETED SUCCESSFULLY !!!!" << std::endl;
  return EXIT_SUCCESS;
}
