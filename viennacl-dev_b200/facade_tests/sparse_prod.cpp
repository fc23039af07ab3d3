// sparse_prod.cpp -- the sparse matrix-vector product tests of the reference (tests/src/sparse.cpp:130-260, 387-427,
// 830-871; tests/src/self_assign.cpp:310-360) re-stated against the B200 facade: same user code (copy / prod / project /
// += / -= / aliasing), same error metric (tests/src/sparse.cpp:66-101) and the same epsilon (1e-12 for double, :1113).
// The fixture mat65k.mtx is not shipped with the reference, so the matrix is synthetic: a banded random pattern with the
// 0.5 diagonal the reference adds (:343-346, "get rid of round-off errors by making row-sums unequal to zero"), plus a few very long rows to reach the whole-CTA row path.
// Checker arithmetic is the plain STL loop the reference test uses for its own expected values (:389, prod(std_matrix, rhs)).
#include <cstdlib>
#include <cmath>
#include <iostream>
#include <map>
#include <vector>

#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/sliced_ell_matrix.hpp"
#include "viennacl/ell_matrix.hpp"
#include "viennacl/hyb_matrix.hpp"
#include "viennacl/coordinate_matrix.hpp"
#include "viennacl/linalg/prod.hpp"
#include "viennacl/linalg/inner_prod.hpp"
#include "viennacl/linalg/norm_2.hpp"
#include "viennacl/tools/matrix_generation.hpp"

// built twice, like the reference's test (tests/src/sparse.cpp:1098-1140 runs float, eps 1e-4, then double, eps 1e-12):
// sparse_prod (double) and sparse_prod_float (-DNUMERIC_T=float)
#ifndef NUMERIC_T
#define NUMERIC_T double
#endif
typedef NUMERIC_T NumericT;
typedef std::vector< std::map<unsigned int, NumericT> > StlMatrix;

static unsigned long long rng_state = 88172645463325252ULL;
static NumericT randomNumber()
{
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return NumericT(rng_state >> 11) / NumericT(1ULL << 53);
}

static std::vector<NumericT> stl_prod(StlMatrix const & A, std::vector<NumericT> const & x)
{
  std::vector<NumericT> y(A.size());
  for (std::size_t i = 0; i < A.size(); ++i)
  {
    NumericT val = 0;
    for (std::map<unsigned int, NumericT>::const_iterator it = A[i].begin(); it != A[i].end(); ++it) val += it->second * x[it->first];
    y[i] = val;
  }
  return y;
}

// tests/src/sparse.cpp:66-101
static NumericT diff(std::vector<NumericT> const & v1, viennacl::vector<NumericT> const & v2)
{
  std::vector<NumericT> v2_cpu(v2.size());
  viennacl::backend::finish();
  viennacl::copy(v2.begin(), v2.end(), v2_cpu.begin());
  NumericT norm_inf = 0;
  for (std::size_t i = 0; i < v1.size(); ++i)
  {
    NumericT m = std::max(std::fabs(v2_cpu[i]), std::fabs(v1[i]));
    NumericT d = m > 0 ? std::fabs(v2_cpu[i] - v1[i]) / m : 0;
    if (d > 0.0001)
    {
      std::cout << "Error at entry " << i << ": Should: " << v1[i] << " vs. Is: " << v2_cpu[i] << std::endl;
      std::exit(EXIT_FAILURE);
    }
    norm_inf = std::max(norm_inf, d);
  }
  return norm_inf;
}

#define CHECK(what, expected, got)                                                                   \
  do { NumericT e_ = std::fabs(diff(expected, got));                                                 \
       if (!(e_ <= epsilon)) { std::cout << "# Error at operation: " << what << "\n  diff: " << e_ << std::endl; return EXIT_FAILURE; } \
       std::cout << "  ok  " << what << "  (diff " << e_ << ")" << std::endl; } while (0)

// tests/src/sparse.cpp:130-260
template<typename VCL_MatrixT>
static int strided_matrix_vector_product_test(NumericT epsilon, std::vector<NumericT> & result, std::vector<NumericT> const & rhs,
                                              viennacl::vector<NumericT> & vcl_result, viennacl::vector<NumericT> & vcl_rhs)
{
  StlMatrix std_A(5);
  std_A[0][0] = 2.0; std_A[0][2] = -1.0;
  std_A[1][0] = 3.0; std_A[1][2] = -5.0;
  std_A[2][1] = 5.0; std_A[2][2] = -2.0;
  std_A[3][2] = 1.0; std_A[3][3] = -6.0;
  std_A[4][1] = 7.0; std_A[4][2] = -5.0;
  // project(result, slice(1, 3, 5)) = prod(std_A, project(rhs, slice(3, 2, 4)))
  for (std::size_t i = 0; i < 5; ++i)
  {
    NumericT val = 0;
    for (std::map<unsigned int, NumericT>::const_iterator it = std_A[i].begin(); it != std_A[i].end(); ++it) val += it->second * rhs[3 + 2 * it->first];
    result[1 + 3 * i] = val;
  }
  VCL_MatrixT vcl_sparse_matrix2;
  viennacl::copy(std_A, vcl_sparse_matrix2);
  viennacl::vector<NumericT> vec(4);
  vec(0) = rhs[3]; vec(1) = rhs[5]; vec(2) = rhs[7]; vec(3) = rhs[9];

  viennacl::project(vcl_result, viennacl::slice(1, 3, 5)) = viennacl::linalg::prod(vcl_sparse_matrix2, viennacl::project(vcl_rhs, viennacl::slice(3, 2, 4)));
  CHECK("matrix-vector product with strided vectors, part 1", result, vcl_result);

  vcl_result(1) = 1.0; vcl_result(4) = 1.0; vcl_result(7) = 1.0; vcl_result(10) = 1.0; vcl_result(13) = 1.0;
  viennacl::project(vcl_result, viennacl::slice(1, 3, 5)) = viennacl::linalg::prod(vcl_sparse_matrix2, vec);
  CHECK("matrix-vector product with strided vectors, part 2", result, vcl_result);

  // ranges: project(result, range(2, 7)) = prod(std_A, project(rhs, range(4, 8)))
  for (std::size_t i = 0; i < 5; ++i)
  {
    NumericT val = 0;
    for (std::map<unsigned int, NumericT>::const_iterator it = std_A[i].begin(); it != std_A[i].end(); ++it) val += it->second * rhs[4 + it->first];
    result[2 + i] = val;
  }
  viennacl::project(vcl_result, viennacl::range(2, 7)) = viennacl::linalg::prod(vcl_sparse_matrix2, viennacl::project(vcl_rhs, viennacl::range(4, 8)));
  CHECK("matrix-vector product with ranged vectors", result, vcl_result);
  return EXIT_SUCCESS;
}

template<typename VCL_MatrixT>
static int product_tests(const char *name, NumericT epsilon, StlMatrix const & std_matrix, std::vector<NumericT> const & rhs)
{
  std::cout << "Testing products: " << name << std::endl;
  std::vector<NumericT> result(rhs);
  viennacl::vector<NumericT> vcl_rhs(rhs.size()), vcl_result(rhs.size());
  viennacl::copy(rhs.begin(), rhs.end(), vcl_rhs.begin());
  VCL_MatrixT vcl_matrix;
  viennacl::copy(std_matrix, vcl_matrix);

  result = stl_prod(std_matrix, rhs);
  vcl_result = viennacl::linalg::prod(vcl_matrix, vcl_rhs);
  CHECK("matrix-vector product", result, vcl_result);

  if (strided_matrix_vector_product_test<VCL_MatrixT>(epsilon, result, rhs, vcl_result, vcl_rhs) != EXIT_SUCCESS) return EXIT_FAILURE;

  result = stl_prod(std_matrix, rhs);
  for (std::size_t i = 0; i < result.size(); ++i) result[i] += rhs[i];
  vcl_result = vcl_rhs;
  vcl_result += viennacl::linalg::prod(vcl_matrix, vcl_rhs);
  CHECK("matrix-vector product (+=)", result, vcl_result);

  result = stl_prod(std_matrix, rhs);
  for (std::size_t i = 0; i < result.size(); ++i) result[i] = rhs[i] - result[i];
  vcl_result = vcl_rhs;
  vcl_result -= viennacl::linalg::prod(vcl_matrix, vcl_rhs);
  CHECK("matrix-vector product (-=)", result, vcl_result);

  // aliasing, tests/src/self_assign.cpp:310-360: x = prod(A, x)
  result = stl_prod(std_matrix, rhs);
  vcl_result = vcl_rhs;
  vcl_result = viennacl::linalg::prod(vcl_matrix, vcl_result);
  CHECK("self-assignment x = prod(A, x)", result, vcl_result);

  // constructor from expression
  viennacl::vector<NumericT> vcl_fresh = viennacl::linalg::prod(vcl_matrix, vcl_rhs);
  CHECK("vector constructed from prod(A, x)", result, vcl_fresh);

  // products after clear() (tests/src/sparse.cpp:988-1060): the cleared matrix holds no entries, A x = 0, y += A x leaves y alone
  vcl_matrix.clear();
  result = std::vector<NumericT>(result.size(), NumericT(0));
  vcl_result = vcl_rhs;                                           // stale content must be overwritten
  vcl_result = viennacl::linalg::prod(vcl_matrix, vcl_rhs);
  CHECK("matrix-vector product after clear()", result, vcl_result);
  result = rhs;
  vcl_result = vcl_rhs;
  vcl_result += viennacl::linalg::prod(vcl_matrix, vcl_rhs);
  CHECK("matrix-vector product (+=) after clear()", result, vcl_result);
  return EXIT_SUCCESS;
}

int main()
{
  const NumericT epsilon = sizeof(NumericT) == 4 ? NumericT(1e-4) : NumericT(1e-12);   // tests/src/sparse.cpp:1105, :1113
  const std::size_t n = 65025;                        // size of the reference's fixture

  StlMatrix std_matrix(n);
  std::vector<NumericT> rhs(n);
  for (std::size_t i = 0; i < n; ++i)
  {
    std::size_t per_row = 1 + std::size_t(randomNumber() * 9);
    for (std::size_t k = 0; k < per_row; ++k)
    {
      long j = long(i) + long(randomNumber() * 600.0) - 300;
      if (j < 0) j = 0;
      if (j >= long(n)) j = long(n) - 1;
      std_matrix[i][static_cast<unsigned int>(j)] = NumericT(-0.02) * randomNumber();   // no cancellation in A*x or b - A*x:
                                                                               // the metric below is a RELATIVE error per entry
    }
    std_matrix[i][static_cast<unsigned int>(i)] = 0.5;
    rhs[i] = NumericT(1.0) + randomNumber();
  }
  std_matrix[n - 1][static_cast<unsigned int>(n - 1)] = 0.5;   // last column populated -> copy() infers a square matrix
  for (std::size_t r = 100; r < n; r += 20000)                  // rows longer than one row block (2048) and than 4 KiB staging
    for (std::size_t j = 0; j < 5000; ++j) std_matrix[r][static_cast<unsigned int>((j * 13) % n)] = NumericT(-0.02) * randomNumber();
  std_matrix[7].clear();                                         // an empty row

  if (product_tests< viennacl::compressed_matrix<NumericT> >("compressed_matrix", epsilon, std_matrix, rhs) != EXIT_SUCCESS) return EXIT_FAILURE;
  if (product_tests< viennacl::sliced_ell_matrix<NumericT> >("sliced_ell_matrix", epsilon, std_matrix, rhs) != EXIT_SUCCESS) return EXIT_FAILURE;
  if (product_tests< viennacl::hyb_matrix<NumericT> >("hyb_matrix", epsilon, std_matrix, rhs) != EXIT_SUCCESS) return EXIT_FAILURE;
  if (product_tests< viennacl::coordinate_matrix<NumericT> >("coordinate_matrix", epsilon, std_matrix, rhs) != EXIT_SUCCESS) return EXIT_FAILURE;
  {
    viennacl::coordinate_matrix<NumericT> M;
    viennacl::copy(std_matrix, M);
    StlMatrix back;
    viennacl::copy(M, back);
    bool same = back.size() == std_matrix.size();
    for (std::size_t i = 0; same && i < n; ++i) same = back[i] == std_matrix[i];
    if (!same) { std::cout << "# coordinate_matrix round trip differs" << std::endl; return EXIT_FAILURE; }
    std::cout << "  ok  copy(host -> coordinate_matrix -> host) is the identity" << std::endl;
  }
  {
    // SELL-C-sigma (extension, sigma = 1 in the reference): rows sorted by length inside windows of 1024 rows before slicing;
    // same product, bit for bit, as the sigma = 1 matrix, with less padding
    viennacl::sliced_ell_matrix<NumericT> S1, S2;
    viennacl::copy(std_matrix, S1);
    S2.sigma(1024);
    viennacl::copy(std_matrix, S2);
    viennacl::vector<NumericT> vx(n), y1(n), y2(n);
    viennacl::copy(rhs.begin(), rhs.end(), vx.begin());
    y1 = viennacl::linalg::prod(S1, vx);
    y2 = viennacl::linalg::prod(S2, vx);
    std::vector<NumericT> h1(n), h2(n);
    viennacl::backend::finish();
    viennacl::copy(y1.begin(), y1.end(), h1.begin());
    viennacl::copy(y2.begin(), y2.end(), h2.begin());
    if (h1 != h2 || S2.sigma() != 1024 || S2.padded_nnz() == 0) { std::cout << "# SELL-C-sigma product differs from sigma = 1" << std::endl; return EXIT_FAILURE; }
    std::cout << "  ok  sliced_ell_matrix with sigma = 1024: product identical to sigma = 1 (" << S2.padded_nnz() << " stored entries)" << std::endl;
  }
  {
    // ell_matrix stores rows * (longest row) entries: tests/src/sparse.cpp:805-828 uses the same matrix; here the 5000-entry
    // rows are dropped so that the padded storage stays small
    StlMatrix ell_input(std_matrix);
    for (std::size_t r = 100; r < n; r += 20000) { ell_input[r].clear(); ell_input[r][static_cast<unsigned int>(r)] = 0.5; }
    if (product_tests< viennacl::ell_matrix<NumericT> >("ell_matrix", epsilon, ell_input, rhs) != EXIT_SUCCESS) return EXIT_FAILURE;
    viennacl::ell_matrix<NumericT> E;
    viennacl::copy(ell_input, E);
    StlMatrix back;
    viennacl::copy(E, back);
    bool same = back.size() == ell_input.size();
    for (std::size_t i = 0; same && i < n; ++i) same = back[i] == ell_input[i];
    if (!same) { std::cout << "# ell_matrix round trip differs" << std::endl; return EXIT_FAILURE; }
    std::cout << "  ok  copy(host -> ell_matrix -> host) is the identity (width " << E.maxnnz() << ")" << std::endl;
  }

  // round trip device -> host (tests/src/sparse.cpp:104-127 diff(cpu_A, vcl_A))
  {
    viennacl::compressed_matrix<NumericT> A;
    viennacl::copy(std_matrix, A);
    StlMatrix back;
    viennacl::copy(A, back);
    if (back.size() != std_matrix.size()) { std::cout << "# round trip: wrong row count" << std::endl; return EXIT_FAILURE; }
    for (std::size_t i = 0; i < n; ++i)
      if (back[i] != std_matrix[i]) { std::cout << "# round trip: row " << i << " differs" << std::endl; return EXIT_FAILURE; }
    std::cout << "  ok  copy(host -> compressed_matrix -> host) is the identity (" << A.nnz() << " nonzeros, " << A.blocks1() << " row blocks)" << std::endl;
  }

  // BLAS-1 helpers the drivers use
  {
    viennacl::vector<NumericT> a(n), b(n);
    viennacl::copy(rhs.begin(), rhs.end(), a.begin());
    b = NumericT(2) * a;
    NumericT ip = viennacl::linalg::inner_prod(a, b), nrm = viennacl::linalg::norm_2(a);
    double ip_ref = 0; for (std::size_t i = 0; i < n; ++i) ip_ref += double(rhs[i]) * 2 * double(rhs[i]);
    const double blas_eps = sizeof(NumericT) == 4 ? 1e-5 : 1e-10;
    if (std::fabs(ip - ip_ref) > blas_eps * ip_ref || std::fabs(double(nrm) * nrm * 2 - ip_ref) > blas_eps * ip_ref)
    { std::cout << "# inner_prod / norm_2 mismatch: " << ip << " " << nrm << " vs " << ip_ref << std::endl; return EXIT_FAILURE; }
    viennacl::vector<NumericT> c = a - b;            // c = -a
    c += a;                                          // c = 0
    if (NumericT(viennacl::linalg::norm_2(c)) != 0) { std::cout << "# vector expression mismatch" << std::endl; return EXIT_FAILURE; }
    std::cout << "  ok  inner_prod / norm_2 / vector expressions" << std::endl;
  }

  // generator (tools/matrix_generation.hpp:47-88)
  {
    viennacl::compressed_matrix<NumericT> L;
    viennacl::tools::generate_fdm_laplace(L, 31, 17);
    StlMatrix host;
    viennacl::copy(L, host);
    bool ok = host.size() == 31 * 17 && L.nnz() == 5 * 31 * 17 - 2 * 31 - 2 * 17;
    for (std::size_t i = 0; ok && i < host.size(); ++i) ok = host[i][static_cast<unsigned int>(i)] == 4.0;
    if (!ok) { std::cout << "# generate_fdm_laplace mismatch" << std::endl; return EXIT_FAILURE; }
    std::cout << "  ok  generate_fdm_laplace" << std::endl;
  }

  std::cout << "!!!! TEST COMPLETED SUCCESSFULLY !!!!" << std::endl;
  return EXIT_SUCCESS;
}
