// matrix_market.cpp -- the ingest step of the reference's tests and tutorials (read_matrix_market_file into an STL matrix,
// copy to the device, prod / solve; tests/src/sparse.cpp:330-336, examples/tutorial/iterative.cpp:77-81) against the B200
// facade.  The file is written by this program (the reference's fixture mat65k.mtx is not shipped): general, symmetric
// and pattern headers, comments, 0- and 1-based indices, malformed input.  The HOST part (parsing) is checked even without
// a device when the program is started with the argument "host-only".
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/linalg/prod.hpp"
#include "viennacl/linalg/norm_2.hpp"
#include "viennacl/linalg/cg.hpp"
#include "viennacl/io/matrix_market.hpp"

typedef double ScalarType;
typedef std::vector< std::map<unsigned int, ScalarType> > StlMatrix;

static int failures = 0;
static void expect(bool ok, const char *what)
{
  std::cout << (ok ? "  ok  " : "# FAILED: ") << what << std::endl;
  if (!ok) ++failures;
}

int main(int argc, char **argv)
{
  const bool host_only = argc > 1 && std::string(argv[1]) == "host-only";
  const std::string dir = (std::getenv("TMPDIR") ? std::string(std::getenv("TMPDIR")) : std::string("/tmp")) + "/";
  const std::string f_general = dir + "vcl_b200_general.mtx", f_sym = dir + "vcl_b200_sym.mtx", f_pat = dir + "vcl_b200_pattern.mtx",
                    f_bad = dir + "vcl_b200_bad.mtx";

  // a 2-D Laplacian 30 x 20, written through write_matrix_market_file
  const std::size_t nx = 30, ny = 20, n = nx * ny;
  StlMatrix A(n);
  for (std::size_t j = 0; j < ny; ++j)
    for (std::size_t i = 0; i < nx; ++i)
    {
      const unsigned int r = static_cast<unsigned int>(i + nx * j);
      A[r][r] = 4.0 + 1e-3 * r;                           // non-trivial values exercise the 17-digit round trip
      if (i > 0) A[r][r - 1] = -1.0;
      if (i + 1 < nx) A[r][r + 1] = -1.0;
      if (j > 0) A[r][r - static_cast<unsigned int>(nx)] = -1.0;
      if (j + 1 < ny) A[r][r + static_cast<unsigned int>(nx)] = -1.0;
    }
  viennacl::io::write_matrix_market_file(A, f_general);
  StlMatrix B;
  long lines = viennacl::io::read_matrix_market_file(B, f_general);
  expect(lines > 2 && B == A, "write_matrix_market_file -> read_matrix_market_file is the identity (general, 1-based)");

  {
    std::ofstream w(f_sym.c_str());                        // lower triangle only, with comments and a blank line
    w << "%%MatrixMarket matrix coordinate real symmetric\n% a comment\n\n" << n << " " << n << " " << (n + (nx - 1) * ny + nx * (ny - 1)) << "\n";
    w.precision(17);
    for (std::size_t r = 0; r < n; ++r)
      for (std::map<unsigned int, ScalarType>::const_iterator it = A[r].begin(); it != A[r].end(); ++it)
        if (it->first <= r) w << r + 1 << " " << it->first + 1 << " " << it->second << "\n";
  }
  StlMatrix S;
  expect(viennacl::io::read_matrix_market_file(S, f_sym) > 0 && S == A, "symmetric header mirrors the off-diagonal entries");

  {
    std::ofstream w(f_pat.c_str());
    w << "%%MatrixMarket matrix coordinate pattern general\n3 4 3\n0 0\n1 3\n2 1\n";
  }
  StlMatrix P;
  expect(viennacl::io::read_matrix_market_file(P, f_pat, 0) > 0 && P.size() == 3 && P[0][0] == 1.0 && P[1][3] == 1.0 && P[2][1] == 1.0 &&
         P[0].size() + P[1].size() + P[2].size() == 3, "pattern matrix with index_base 0 gets unit values");

  {
    std::ofstream w(f_bad.c_str());
    w << "%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n3 1 2.0\n";
  }
  StlMatrix Bad;
  std::cerr << "(the next two error messages are expected)" << std::endl;
  expect(viennacl::io::read_matrix_market_file(Bad, f_bad) == 0, "out-of-range index is rejected (returns 0)");
  expect(viennacl::io::read_matrix_market_file(Bad, dir + "vcl_b200_does_not_exist.mtx") == EXIT_FAILURE, "missing file returns EXIT_FAILURE");

  if (!host_only)
  {
    // iterative.cpp:77-126: read, copy to the device, solve
    viennacl::compressed_matrix<ScalarType> vcl_A;
    expect(viennacl::io::read_matrix_market_file(vcl_A, f_sym) > 0 && vcl_A.size1() == n && vcl_A.size2() == n && vcl_A.nnz() == 5 * n - 2 * nx - 2 * ny,
           "read_matrix_market_file(compressed_matrix &) uploads the parsed matrix");
    viennacl::vector<ScalarType> rhs = viennacl::scalar_vector<ScalarType>(n, 1.0);
    viennacl::linalg::cg_tag tag(1e-10, 1000);
    viennacl::vector<ScalarType> x = viennacl::linalg::solve(vcl_A, rhs, tag);
    viennacl::vector<ScalarType> r = rhs - viennacl::linalg::prod(vcl_A, x);
    expect(ScalarType(viennacl::linalg::norm_2(r)) < 1e-8 * ScalarType(viennacl::linalg::norm_2(rhs)), "CG on the matrix read from the file");
  }
  std::remove(f_general.c_str()); std::remove(f_sym.c_str()); std::remove(f_pat.c_str()); std::remove(f_bad.c_str());
  if (failures) { std::cout << failures << " check(s) FAILED" << std::endl; return EXIT_FAILURE; }
  std::cout << "!!!! TEST COMPLETED SUCCESSFULLY !!!!" << std::endl;
  return EXIT_SUCCESS;
}
