// ref_tree_test.cu -- the UNMODIFIED reference tree (/root/reference: its vector / compressed_matrix / sliced_ell_matrix classes, copy(),
// BLAS-1 CUDA kernels and its OWN solver drivers cg.hpp:128-187, bicgstab.hpp:97-215, gmres.hpp:181-367) with only the CUDA_MEMORY arms
// of the hot path re-bound to libvcl_b200.so (viennacl/linalg/b200/binding.hpp; dispatch headers patched by sed at build time).
// Prints one line per check; tests/test_ref_binding.py compares the counts with the reference's host-backend goldens.
#include <cmath>
#include <cstdio>
#include <ctime>
#include <string>
#include <map>
#include <vector>
#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/sliced_ell_matrix.hpp"
#include "viennacl/linalg/prod.hpp"
#include "viennacl/linalg/norm_2.hpp"
#include "viennacl/linalg/cg.hpp"
#include "viennacl/linalg/bicgstab.hpp"
#include "viennacl/linalg/gmres.hpp"

typedef std::vector<std::map<unsigned int, double> > HostMatrix;

// 2-D 5-point stencil with first-order upwind convection (the generator of oracle/vcl_oracle.c, restated): diag 4 + cx + cy
static HostMatrix stencil2d(int nx, int ny, double cx, double cy)
{
  HostMatrix A((size_t)nx * ny);
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i)
    {
      const unsigned int r = (unsigned int)(j * nx + i);
      A[r][r] = 4.0 + cx + cy;
      if (i > 0) A[r][r - 1] = -1.0 - cx;
      if (i < nx - 1) A[r][r + 1] = -1.0;
      if (j > 0) A[r][r - nx] = -1.0 - cy;
      if (j < ny - 1) A[r][r + nx] = -1.0;
    }
  return A;
}

static double max_rel(std::vector<double> const & a, std::vector<double> const & b)
{
  double m = 0;
  for (size_t i = 0; i < a.size(); ++i)
  {
    const double d = std::fabs(a[i] - b[i]), s = std::max(std::fabs(a[i]), std::fabs(b[i]));
    if (s > 0) m = std::max(m, d / s);
  }
  return m;
}

template<class MatrixT>
static void run(const char *name, const char *fmt, HostMatrix const & H, bool spd)
{
  const size_t n = H.size();
  MatrixT A;
  viennacl::copy(H, A);                                            // the reference's own copy(): CUDA memory is the default domain
  std::vector<double> hx(n), hy(n, 0.0), ref(n, 0.0);
  for (size_t i = 0; i < n; ++i) hx[i] = 1.0 + double((i * 2654435761u) % 1000u) / 1000.0;
  for (size_t r = 0; r < n; ++r)
    for (std::map<unsigned int, double>::const_iterator it = H[r].begin(); it != H[r].end(); ++it) ref[r] += it->second * hx[it->first];
  viennacl::vector<double> x(n), y(n), b = viennacl::scalar_vector<double>(n, 1.0);
  viennacl::copy(hx, x);
  y = viennacl::linalg::prod(A, x);                                // -> b200::prod_impl
  viennacl::copy(y, hy);
  std::printf("REFBIND %s %s prod max_rel_err %.3e\n", name, fmt, max_rel(hy, ref));
  y += viennacl::linalg::prod(A, x);                               // the += form (alpha = 1, beta = 1)
  viennacl::copy(y, hy);
  for (size_t i = 0; i < n; ++i) ref[i] *= 2.0;
  std::printf("REFBIND %s %s prod_pluseq max_rel_err %.3e\n", name, fmt, max_rel(hy, ref));

  if (spd)
  {
    viennacl::linalg::cg_tag tag(1e-8, 1000);
    viennacl::vector<double> sol = viennacl::linalg::solve(A, b, tag);          // the reference's pipelined CG driver
    viennacl::vector<double> res = viennacl::linalg::prod(A, sol); res -= b;
    std::printf("REFBIND %s %s cg iters %u error %.6e true_residual %.3e\n", name, fmt, (unsigned)tag.iters(), tag.error(),
                viennacl::linalg::norm_2(res) / viennacl::linalg::norm_2(b));
  }
  {
    viennacl::linalg::bicgstab_tag tag(1e-8, 1000);
    viennacl::vector<double> sol = viennacl::linalg::solve(A, b, tag);          // the reference's pipelined BiCGStab driver
    viennacl::vector<double> res = viennacl::linalg::prod(A, sol); res -= b;
    std::printf("REFBIND %s %s bicgstab iters %u error %.6e true_residual %.3e\n", name, fmt, (unsigned)tag.iters(), tag.error(),
                viennacl::linalg::norm_2(res) / viennacl::linalg::norm_2(b));
  }
  {
    viennacl::linalg::gmres_tag tag(1e-8, 1000, 30);
    viennacl::vector<double> sol = viennacl::linalg::solve(A, b, tag);          // the reference's pipelined GMRES driver
    viennacl::vector<double> res = viennacl::linalg::prod(A, sol); res -= b;
    std::printf("REFBIND %s %s gmres iters %u error %.6e true_residual %.3e\n", name, fmt, (unsigned)tag.iters(), tag.error(),
                viennacl::linalg::norm_2(res) / viennacl::linalg::norm_2(b));
  }
}

// `bench`: BASELINE configs[0] (CG, 2-D Laplacian 1024^2, tol 1e-8) through the reference's own driver, timed like
// examples/benchmarks/solver.cpp:106-120 (finish() on both sides).  Built twice: with the binding (new kernels under the reference's
// driver, which still reads its 768 partial sums back every iteration, cg.hpp:168) and without (the reference's own CUDA kernels).
static void bench()
{
  const HostMatrix H = stencil2d(1024, 1024, 0.0, 0.0);
  viennacl::compressed_matrix<double> A;
  viennacl::copy(H, A);
  viennacl::vector<double> b = viennacl::scalar_vector<double>(H.size(), 1.0);
  for (int rep = 0; rep < 3; ++rep)
  {
    viennacl::linalg::cg_tag tag(1e-8, 5000);
    viennacl::backend::finish();
    const double t0 = (double)clock() / CLOCKS_PER_SEC;
    struct timespec a, c; clock_gettime(CLOCK_MONOTONIC, &a);
    viennacl::vector<double> sol = viennacl::linalg::solve(A, b, tag);
    viennacl::backend::finish();
    clock_gettime(CLOCK_MONOTONIC, &c);
    const double sec = (c.tv_sec - a.tv_sec) + 1e-9 * (c.tv_nsec - a.tv_nsec);
    (void)t0;
    std::printf("REFBIND bench cg_lap2d_1024 rep %d iters %u error %.6e seconds %.6f iterations_per_sec %.1f\n", rep, (unsigned)tag.iters(), tag.error(), sec,
                tag.iters() / sec);
  }
}

int main(int argc, char **argv)
{
  try
  {
    if (argc > 1 && std::string(argv[1]) == "bench") { bench(); std::printf("REFBIND DONE\n"); return 0; }
    const HostMatrix L = stencil2d(63, 65, 0.0, 0.0), C = stencil2d(48, 50, 0.5, 0.0);
    run<viennacl::compressed_matrix<double> >("lap2d_63x65", "csr", L, true);
    run<viennacl::compressed_matrix<double> >("cd2d_48x50", "csr", C, false);
    run<viennacl::sliced_ell_matrix<double> >("lap2d_63x65", "sell", L, true);
    run<viennacl::sliced_ell_matrix<double> >("cd2d_48x50", "sell", C, false);
#ifdef VCL_REF_BINDING
    long long launches = 0;
    ViennaCLBackendLaunchCount(viennacl::linalg::b200::backend(), &launches);
    std::printf("REFBIND launches_of_libvcl_b200 %lld\n", launches);
#endif
    std::printf("REFBIND DONE\n");
  }
  catch (std::exception const & e)
  {
    std::printf("REFBIND EXCEPTION %s\n", e.what());
    return 1;
  }
  return 0;
}
