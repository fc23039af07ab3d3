// wrap_cuda_buffer.cpp -- examples/tutorial/wrap-cuda-buffer.cu:60-110 and the CSR wrap constructor
// (compressed_matrix.hpp:740-781) against the B200 facade: user-owned device buffers are wrapped without a copy, used in
// prod() / solve(), and stay valid (and un-freed) after the ViennaCL objects are gone.  The buffers come from the C-ABI
// allocator so that this program needs no CUDA toolkit on the host side; any cudaMalloc'ed pointer works the same way.
#include <cstdlib>
#include <cmath>
#include <iostream>
#include <vector>

#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/linalg/prod.hpp"
#include "viennacl/linalg/norm_2.hpp"
#include "viennacl/linalg/cg.hpp"

typedef double ScalarType;

template<typename T>
static T *device_array(std::vector<T> const & host)
{
  void *p = NULL;
  ViennaCLBackend h = viennacl::backend::b200::handle();
  viennacl::backend::b200::check(ViennaCLCUDAMemAlloc(h, &p, host.size() * sizeof(T)));
  viennacl::backend::b200::check(ViennaCLCUDAMemWrite(h, p, 0, &host[0], host.size() * sizeof(T), 0));
  return static_cast<T*>(p);
}

template<typename T>
static std::vector<T> host_array(const T *dev, std::size_t n)
{
  std::vector<T> host(n);
  viennacl::backend::b200::check(ViennaCLCUDAMemRead(viennacl::backend::b200::handle(), dev, 0, &host[0], n * sizeof(T), 0));
  return host;
}

int main()
{
  // Part 1: vectors (wrap-cuda-buffer.cu:60-105)
  std::size_t size = 10;
  ScalarType *cuda_x = device_array(std::vector<ScalarType>(size, 1.0));
  ScalarType *cuda_y = device_array(std::vector<ScalarType>(size, 2.0));
  {
    viennacl::vector<ScalarType> vcl_vec1(cuda_x, viennacl::CUDA_MEMORY, size);
    viennacl::vector<ScalarType> vcl_vec2(cuda_y, viennacl::CUDA_MEMORY, size);
    vcl_vec1 = viennacl::scalar_vector<ScalarType>(size, ScalarType(1.0));
    vcl_vec2 = viennacl::scalar_vector<ScalarType>(size, ScalarType(2.0));
    vcl_vec1 += vcl_vec2;
    std::cout << "Result with ViennaCL: " << vcl_vec1 << std::endl;
  }
  std::vector<ScalarType> after = host_array(cuda_x, size);      // the buffer outlives the wrapper and holds the result
  for (std::size_t i = 0; i < size; ++i)
    if (after[i] != 3.0) { std::cout << "# wrapped vector: entry " << i << " is " << after[i] << std::endl; return EXIT_FAILURE; }

  // Part 2: CSR arrays of a 1-D Laplacian, wrapped (compressed_matrix.hpp:740-781), then prod() and solve()
  const unsigned int n = 1000;
  std::vector<unsigned int> rp(n + 1), ci;
  std::vector<ScalarType> va;
  for (unsigned int i = 0; i < n; ++i)
  {
    rp[i] = static_cast<unsigned int>(ci.size());
    if (i > 0) { ci.push_back(i - 1); va.push_back(-1.0); }
    ci.push_back(i); va.push_back(2.0);
    if (i + 1 < n) { ci.push_back(i + 1); va.push_back(-1.0); }
  }
  rp[n] = static_cast<unsigned int>(ci.size());
  unsigned int *d_rp = device_array(rp), *d_ci = device_array(ci);
  ScalarType *d_va = device_array(va);
  ScalarType *d_b = device_array(std::vector<ScalarType>(n, 1.0));
  ScalarType *d_out = device_array(std::vector<ScalarType>(n, -7.0));
  {
    viennacl::compressed_matrix<ScalarType> A(d_rp, d_ci, d_va, viennacl::CUDA_MEMORY, n, n, ci.size());
    viennacl::vector<ScalarType> b(d_b, viennacl::CUDA_MEMORY, n), out(d_out, viennacl::CUDA_MEMORY, n);
    out = viennacl::linalg::prod(A, b);                           // written straight into the user's buffer
    viennacl::linalg::cg_tag tag(1e-10, 2000);
    viennacl::vector<ScalarType> x = viennacl::linalg::solve(A, b, tag);
    viennacl::vector<ScalarType> r = b - viennacl::linalg::prod(A, x);
    ScalarType res = ScalarType(viennacl::linalg::norm_2(r)) / ScalarType(viennacl::linalg::norm_2(b));
    std::cout << "CG on wrapped CSR: " << tag.iters() << " iterations, true relative residual " << res << std::endl;
    if (!(res < 1e-8)) { std::cout << "# solve on wrapped CSR failed" << std::endl; return EXIT_FAILURE; }
  }
  std::vector<ScalarType> y = host_array(d_out, n);
  for (unsigned int i = 0; i < n; ++i)
  {
    ScalarType expect = (i == 0 || i == n - 1) ? 1.0 : 0.0;
    if (y[i] != expect) { std::cout << "# prod into wrapped buffer: entry " << i << " is " << y[i] << std::endl; return EXIT_FAILURE; }
  }
  if (host_array(d_va, va.size()) != va || host_array(d_ci, ci.size()) != ci) { std::cout << "# wrapped CSR arrays were modified" << std::endl; return EXIT_FAILURE; }

  // the user still owns everything
  ViennaCLBackend h = viennacl::backend::b200::handle();
  ViennaCLCUDAMemFree(h, cuda_x); ViennaCLCUDAMemFree(h, cuda_y); ViennaCLCUDAMemFree(h, d_rp); ViennaCLCUDAMemFree(h, d_ci);
  ViennaCLCUDAMemFree(h, d_va); ViennaCLCUDAMemFree(h, d_b); ViennaCLCUDAMemFree(h, d_out);
  std::cout << "!!!! TUTORIAL COMPLETED SUCCESSFULLY !!!!" << std::endl;
  return EXIT_SUCCESS;
}
