// ref_generic_counts.cpp -- CHECKER (test infrastructure): runs the UNMODIFIED reference (host backend) on the systems of
// viennacl-dev_b200/facade_tests/matrix_free.cpp and prints the iteration counts of its GENERIC solver paths
// (cg.hpp:257-322, bicgstab.hpp:398-489, gmres.hpp:449-631).  The numbers are pinned in tests/golden/generic_solver_counts.json
// (regenerate: make -C oracle generic_counts).  Build: g++ -O2 -I/root/reference oracle/ref_generic_counts.cpp
#include <cstdio>
#include <map>
#include <vector>
#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/linalg/prod.hpp"
#include "viennacl/linalg/jacobi_precond.hpp"
#include "viennacl/linalg/row_scaling.hpp"
#include "viennacl/linalg/cg.hpp"
#include "viennacl/linalg/bicgstab.hpp"
#include "viennacl/linalg/gmres.hpp"

typedef double T;
typedef std::vector< std::map<unsigned int, T> > Stl;

static Stl stencil(std::size_t nx, std::size_t ny, std::size_t nz, T cx, T cy, T cz)
{
  Stl A(nx * ny * nz);
  for (std::size_t k = 0; k < nz; ++k) for (std::size_t j = 0; j < ny; ++j) for (std::size_t i = 0; i < nx; ++i)
  {
    unsigned int r = static_cast<unsigned int>(i + nx * (j + ny * k));
    A[r][r] = (nz > 1 ? 6.0 : 4.0) + cx + cy + (nz > 1 ? cz : 0.0);
    if (i > 0) A[r][r - 1] = -1.0 - cx;
    if (i + 1 < nx) A[r][r + 1] = -1.0;
    if (j > 0) A[r][r - static_cast<unsigned int>(nx)] = -1.0 - cy;
    if (j + 1 < ny) A[r][r + static_cast<unsigned int>(nx)] = -1.0;
    if (nz > 1 && k > 0) A[r][r - static_cast<unsigned int>(nx * ny)] = -1.0 - cz;
    if (nz > 1 && k + 1 < nz) A[r][r + static_cast<unsigned int>(nx * ny)] = -1.0;
  }
  return A;
}

struct ShiftedOp            // y = (L + sigma I) x
{
  ShiftedOp(viennacl::compressed_matrix<T> const & L, T s) : L_(&L), s_(s) {}
  void apply(viennacl::vector_base<T> const & x, viennacl::vector_base<T> & y) const { y = viennacl::linalg::prod(*L_, x); y += s_ * x; }
  std::size_t size1() const { return L_->size1(); }
  viennacl::compressed_matrix<T> const *L_; T s_;
};

int main()
{
  viennacl::context host(viennacl::MAIN_MEMORY);
  printf("{\n");
  {
    Stl sL = stencil(48, 40, 1, 0, 0, 0);
    viennacl::compressed_matrix<T> L(sL.size(), sL.size(), host);
    viennacl::copy(sL, L);
    ShiftedOp op(L, 0.25);
    viennacl::vector<T> rhs = viennacl::scalar_vector<T>(op.size1(), T(-1), host);
    viennacl::linalg::cg_tag t1(1e-9, 1000); viennacl::vector<T> x1 = viennacl::linalg::solve(op, rhs, t1);
    viennacl::linalg::bicgstab_tag t2(1e-9, 1000); viennacl::vector<T> x2 = viennacl::linalg::solve(op, rhs, t2);
    viennacl::linalg::gmres_tag t3(1e-9, 600, 30); viennacl::vector<T> x3 = viennacl::linalg::solve(op, rhs, t3);
    printf(" \"op_cg\": %u, \"op_bicgstab\": %u, \"op_gmres30\": %u,\n", t1.iters(), t2.iters(), t3.iters());
  }
  {
    const std::size_t nx = 40, ny = 36, n = nx * ny;
    Stl sA = stencil(nx, ny, 1, 0, 0, 0);
    for (unsigned int r = 0; r < n; ++r) sA[r][r] = 4.0 + 1.0 + 999.0 * T((r * 7919u) % 1000u) / 999.0;
    viennacl::compressed_matrix<T> A(n, n, host);
    viennacl::copy(sA, A);
    viennacl::vector<T> b = viennacl::scalar_vector<T>(n, T(1), host);
    viennacl::linalg::jacobi_precond< viennacl::compressed_matrix<T> > jac(A, viennacl::linalg::jacobi_tag());
    viennacl::linalg::cg_tag t0(1e-10, 2000); viennacl::vector<T> x0 = viennacl::linalg::solve(A, b, t0);
    viennacl::linalg::cg_tag t1(1e-10, 2000); viennacl::vector<T> x1 = viennacl::linalg::solve(A, b, t1, jac);
    viennacl::linalg::row_scaling< viennacl::compressed_matrix<T> > rs2(A, viennacl::linalg::row_scaling_tag(2));
    viennacl::linalg::cg_tag t2(1e-10, 2000); viennacl::vector<T> x2 = viennacl::linalg::solve(A, b, t2, rs2);
    printf(" \"vardiag_cg\": %u, \"vardiag_cg_jacobi\": %u, \"vardiag_cg_rowscaling2\": %u,\n", t0.iters(), t1.iters(), t2.iters());
  }
  {
    Stl sC = stencil(20, 18, 16, 0.5, 0.25, 0.125);
    viennacl::compressed_matrix<T> C(sC.size(), sC.size(), host);
    viennacl::copy(sC, C);
    viennacl::vector<T> c = viennacl::scalar_vector<T>(sC.size(), T(1), host);
    viennacl::linalg::jacobi_precond< viennacl::compressed_matrix<T> > jac(C, viennacl::linalg::jacobi_tag());
    viennacl::linalg::bicgstab_tag t1(1e-9, 1000); viennacl::vector<T> x1 = viennacl::linalg::solve(C, c, t1, jac);
    viennacl::linalg::gmres_tag t2(1e-9, 600, 20); viennacl::vector<T> x2 = viennacl::linalg::solve(C, c, t2, jac);
    struct ident { void apply(viennacl::vector<T> &) const {} } id;
    viennacl::linalg::gmres_tag t3(1e-9, 600, 20); viennacl::vector<T> x3 = viennacl::linalg::solve(C, c, t3, id);   // Householder path, no preconditioning
    viennacl::linalg::row_scaling< viennacl::compressed_matrix<T> > rs1(C, viennacl::linalg::row_scaling_tag(1));
    viennacl::linalg::row_scaling< viennacl::compressed_matrix<T> > rs0(C, viennacl::linalg::row_scaling_tag(0));
    viennacl::linalg::bicgstab_tag t4(1e-9, 1000); viennacl::vector<T> x4 = viennacl::linalg::solve(C, c, t4, rs1);
    viennacl::linalg::gmres_tag t5(1e-9, 600, 20); viennacl::vector<T> x5 = viennacl::linalg::solve(C, c, t5, rs0);
    printf(" \"cd3d_bicgstab_jacobi\": %u, \"cd3d_gmres20_jacobi\": %u, \"cd3d_gmres20_householder\": %u,\n", t1.iters(), t2.iters(), t3.iters());
    printf(" \"cd3d_bicgstab_rowscaling1\": %u, \"cd3d_gmres20_rowscaling0\": %u\n", (unsigned)t4.iters(), (unsigned)t5.iters());
  }
  printf("}\n");
  return 0;
}
