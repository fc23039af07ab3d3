// api_surface.cpp -- the remaining entries of SURVEY Appendix C (facade API the reference's tests / tutorials touch) that the
// other programs do not exercise: unit_vector, async_copy, result_of, compressed_matrix::operator() / set_entry / resize /
// switch_memory_context, tools::sparse_matrix_adapter + copy(), tools::uniform_random_numbers, the STL overloads of solve().
#include <cstdlib>
#include <cmath>
#include <iostream>
#include <map>
#include <vector>

#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/meta/result_of.hpp"
#include "viennacl/tools/adapter.hpp"
#include "viennacl/tools/random.hpp"
#include "viennacl/tools/matrix_generation.hpp"
#include "viennacl/linalg/prod.hpp"
#include "viennacl/linalg/norm_2.hpp"
#include "viennacl/linalg/inner_prod.hpp"
#include "viennacl/linalg/cg.hpp"
#include "viennacl/linalg/bicgstab.hpp"
#include "viennacl/linalg/gmres.hpp"

typedef double T;
static int failures = 0;
static void expect(bool ok, const char *what)
{
  std::cout << (ok ? "  ok  " : "# FAILED: ") << what << std::endl;
  if (!ok) ++failures;
}

template<typename A, typename B> struct same_type { static const bool value = false; };
template<typename A> struct same_type<A, A> { static const bool value = true; };

int main()
{
  // ---- unit_vector / scalar_vector / async_copy ----
  {
    viennacl::vector<T> e = viennacl::unit_vector<T>(100, 37);
    viennacl::vector<T> ones = viennacl::scalar_vector<T>(100, 1.0);
    expect(std::fabs(T(viennacl::linalg::inner_prod(e, ones)) - 1.0) < 1e-15 && std::fabs(T(viennacl::linalg::norm_2(e)) - 1.0) < 1e-15 && T(e[37]) == 1.0 && T(e[36]) == 0.0,
           "unit_vector: one entry equal to 1");
    std::vector<T> h(1000), back(1000, -1.0);
    for (std::size_t i = 0; i < h.size(); ++i) h[i] = T(i) * 0.5;
    viennacl::vector<T> d(1000);
    viennacl::async_copy(h, d);
    viennacl::async_copy(d, back);
    viennacl::backend::finish();
    expect(back == h, "async_copy host -> device -> host (ordered on the backend's stream, finished by backend::finish())");
  }
  // ---- vector expressions whose operands alias the destination (the reference evaluates them through temporaries) ----
  {
    const std::size_t n = 1000;
    std::vector<T> ha(n), hb(n), hx(n), got(n);
    for (std::size_t i = 0; i < n; ++i) { ha[i] = T(i) * 0.25 + 1.0; hb[i] = 3.0 - T(i) * 0.125; hx[i] = T(i % 7) - 2.0; }
    viennacl::vector<T> a(n), b(n), x(n);
    viennacl::copy(ha, a); viennacl::copy(hb, b);
    bool ok = true;
    viennacl::copy(hx, x);
    x = a + b + x;                                       // third operand is the destination
    viennacl::copy(x, got);
    for (std::size_t i = 0; i < n; ++i) ok = ok && std::fabs(got[i] - (ha[i] + hb[i] + hx[i])) <= 1e-13 * (1.0 + std::fabs(got[i]));
    expect(ok, "x = a + b + x");
    ok = true; viennacl::copy(hx, x);
    x = x + 2.0 * a - x;                                 // first and third operand alias
    viennacl::copy(x, got);
    for (std::size_t i = 0; i < n; ++i) ok = ok && std::fabs(got[i] - 2.0 * ha[i]) <= 1e-13 * (1.0 + std::fabs(got[i]));
    expect(ok, "x = x + 2 a - x");
    ok = true; viennacl::copy(hx, x);
    x = x + x + x;                                       // all three alias: temporary
    viennacl::copy(x, got);
    for (std::size_t i = 0; i < n; ++i) ok = ok && std::fabs(got[i] - 3.0 * hx[i]) <= 1e-13 * (1.0 + std::fabs(got[i]));
    expect(ok, "x = x + x + x");
    ok = true; viennacl::copy(hx, x);
    x += a + b - x;                                      // accumulate form: x + a + b - x
    viennacl::copy(x, got);
    for (std::size_t i = 0; i < n; ++i) ok = ok && std::fabs(got[i] - (ha[i] + hb[i])) <= 1e-12 * (1.0 + std::fabs(got[i]));
    expect(ok, "x += a + b - x");
    ok = true; viennacl::copy(hx, x);
    {
      // overlapping views of one buffer: lo = x[0, 600), hi = x[400, 1000):  lo = hi + lo
      viennacl::vector_range<viennacl::vector<T> > lo(x, viennacl::range(0, 600)), hi(x, viennacl::range(400, 1000));
      lo = hi + lo;
    }
    viennacl::copy(x, got);
    for (std::size_t i = 0; i < 600; ++i) ok = ok && std::fabs(got[i] - (hx[i + 400] + hx[i])) <= 1e-13 * (1.0 + std::fabs(got[i]));
    for (std::size_t i = 600; i < n; ++i) ok = ok && got[i] == hx[i];
    expect(ok, "lo = hi + lo on overlapping ranges of one vector");
  }
  // ---- result_of ----
  expect(same_type<viennacl::result_of::cpu_value_type< viennacl::vector<float> >::type, float>::value &&
         same_type<viennacl::result_of::cpu_value_type< viennacl::compressed_matrix<double> >::type, double>::value &&
         same_type<viennacl::result_of::value_type< viennacl::vector<double> >::type, double>::value &&
         same_type<viennacl::result_of::cpu_value_type<double>::type, double>::value, "result_of::cpu_value_type / value_type");
  // ---- compressed_matrix entry access, insertion, resize, memory context ----
  {
    std::vector< std::map<unsigned int, T> > stl(4);
    stl[0][0] = 2.0; stl[0][2] = -1.0; stl[1][1] = 3.0; stl[2][0] = -1.0; stl[2][3] = 4.0; stl[3][3] = 5.0;
    viennacl::compressed_matrix<T> A;
    viennacl::copy(stl, A);
    expect(A(0, 2) == -1.0 && A(1, 1) == 3.0 && A(1, 0) == 0.0 && A(2, 3) == 4.0, "compressed_matrix::operator()(i, j) reads stored and structural-zero entries");
    A.set_entry(1, 1, 7.0);                       // overwrite
    A.set_entry(1, 3, -2.0);                      // insert
    expect(A(1, 1) == 7.0 && A(1, 3) == -2.0 && A.nnz() == 7, "set_entry overwrites / inserts (entry_proxy assignment)");
    viennacl::vector<T> x = viennacl::scalar_vector<T>(4, 1.0);
    viennacl::vector<T> y = viennacl::linalg::prod(A, x);
    expect(T(y[0]) == 1.0 && T(y[1]) == 5.0 && T(y[2]) == 3.0 && T(y[3]) == 5.0, "prod() after an insertion (row blocks regenerated)");
    A.resize(3, 3, true);
    expect(A.size1() == 3 && A.size2() == 3 && A.nnz() == 4 && A(0, 2) == -1.0 && A(2, 0) == -1.0 && A(1, 1) == 7.0, "resize(3, 3, preserve) drops entries outside the new shape");
    A.resize(5, 5, false);
    expect(A.size1() == 5 && A.nnz() == 0, "resize(5, 5, false) empties the matrix");
    A.switch_memory_context(viennacl::context(viennacl::CUDA_MEMORY));
    bool threw = false;
    try { A.switch_memory_context(viennacl::context(viennacl::MAIN_MEMORY)); } catch (viennacl::memory_exception const &) { threw = true; }
    expect(threw, "switch_memory_context(MAIN_MEMORY) is refused: this build has no host backend");
  }
  // ---- tools::sparse_matrix_adapter, uniform_random_numbers ----
  {
    std::vector< std::map<unsigned int, T> > stl(6);
    viennacl::tools::sparse_matrix_adapter<T> adapted(stl, 6, 8);
    viennacl::tools::uniform_random_numbers<T> rnd;
    bool in_range = true;
    for (unsigned int i = 0; i < 6; ++i)
    {
      const T v = rnd();
      in_range = in_range && v >= 0.0 && v < 1.0;
      adapted(i, i) = 2.0 + v;
      if (i > 0) adapted(i, i - 1) = -1.0;
    }
    viennacl::compressed_matrix<T> A;
    viennacl::copy(adapted, A);
    viennacl::tools::const_sparse_matrix_adapter<T> cadapted(stl, 6, 8);
    expect(in_range && A.size1() == 6 && A.size2() == 8 && A.nnz() == 11 && A(3, 2) == -1.0 && A(5, 5) == cadapted(5, 5) && cadapted(0, 7) == 0.0,
           "sparse_matrix_adapter -> copy() keeps the adapter's shape; uniform_random_numbers in [0, 1)");
  }
  // ---- STL overloads of solve() ----
  {
    const unsigned int nx = 30, ny = 25, n = nx * ny;
    std::vector< std::map<unsigned int, T> > stl(n);
    for (unsigned int j = 0; j < ny; ++j)
      for (unsigned int i = 0; i < nx; ++i)
      {
        const unsigned int r = i + nx * j;
        stl[r][r] = 4.5;
        if (i > 0) stl[r][r - 1] = -1.25;            // nonsymmetric: upwind convection
        if (i + 1 < nx) stl[r][r + 1] = -1.0;
        if (j > 0) stl[r][r - nx] = -1.0;
        if (j + 1 < ny) stl[r][r + nx] = -1.0;
      }
    std::vector<T> rhs(n, 1.0);
    std::vector<T> x1 = viennacl::linalg::solve(stl, rhs, viennacl::linalg::bicgstab_tag(1e-10, 500));
    std::vector<T> x2 = viennacl::linalg::solve(stl, rhs, viennacl::linalg::gmres_tag(1e-10, 600, 30), viennacl::linalg::no_precond());
    T res = 0, diff = 0, nrm = 0;
    for (unsigned int r = 0; r < n; ++r)
    {
      T ax = 0;
      for (std::map<unsigned int, T>::const_iterator it = stl[r].begin(); it != stl[r].end(); ++it) ax += it->second * x1[it->first];
      res += (rhs[r] - ax) * (rhs[r] - ax);
      diff += (x1[r] - x2[r]) * (x1[r] - x2[r]); nrm += x1[r] * x1[r];
    }
    expect(std::sqrt(res / n) < 1e-8 && std::sqrt(diff / nrm) < 1e-7, "solve(std::vector<std::map>, std::vector, bicgstab_tag / gmres_tag[, no_precond])");
    for (unsigned int r = 0; r < n; ++r) { if (r % nx > 0) stl[r][r - 1] = -1.0; }      // symmetric again
    std::vector<T> x3 = viennacl::linalg::solve(stl, rhs, viennacl::linalg::cg_tag(1e-10, 500));
    res = 0;
    for (unsigned int r = 0; r < n; ++r)
    {
      T ax = 0;
      for (std::map<unsigned int, T>::const_iterator it = stl[r].begin(); it != stl[r].end(); ++it) ax += it->second * x3[it->first];
      res += (rhs[r] - ax) * (rhs[r] - ax);
    }
    expect(std::sqrt(res / n) < 1e-8, "solve(std::vector<std::map>, std::vector, cg_tag)");
  }

  if (failures) { std::cout << failures << " check(s) FAILED" << std::endl; return EXIT_FAILURE; }
  std::cout << "!!!! TEST COMPLETED SUCCESSFULLY !!!!" << std::endl;
  return EXIT_SUCCESS;
}
