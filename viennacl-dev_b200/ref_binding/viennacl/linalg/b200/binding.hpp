#ifndef VIENNACL_LINALG_B200_BINDING_HPP_
#define VIENNACL_LINALG_B200_BINDING_HPP_
/* viennacl/linalg/b200/binding.hpp -- the ONE new header a ViennaCL maintainer adds to the UNMODIFIED reference tree so that the
 * `case viennacl::CUDA_MEMORY:` arms of the hot path call libvcl_b200.so.  It is compiled here against
 * the reference where it lies (/root/reference) by ref_binding/Makefile, with the dispatch headers
 *     viennacl/linalg/sparse_matrix_operations.hpp:90-121 (compressed_matrix), :1010-1060 (sliced_ell_matrix)
 *     viennacl/linalg/iterative_operations.hpp:58-418
 * patched at build time by `sed` (cuda:: -> b200:: in those arms; copies under _build/patched/, never committed), and the
 * reference's OWN drivers (cg.hpp:128-187, bicgstab.hpp:97-215, gmres.hpp:181-367), vector class, copy() and BLAS-1 kernels run
 * unchanged on top of the new kernels (ref_tree_test.cu).
 *
 * Stream: the reference's CUDA backend launches everything on the legacy default stream and uses blocking cudaMemcpy; the handle is
 * therefore created ON that stream (cudaStreamLegacy), which keeps every library call ordered with the reference's own kernels. */
#include <cuda_runtime.h>
#include "vcl_b200.h"
#include "viennacl/forwards.h"                  // forward declarations of compressed_matrix / sliced_ell_matrix: this header is included
#include "viennacl/vector.hpp"                  // by the dispatch headers BEFORE the matrix classes are complete, so every function
#include "viennacl/linalg/cuda/common.hpp"      // that touches a matrix is a template on a matrix parameter.  cuda_arg<T>(): cuda/common.hpp:39-146
#include "viennacl/linalg/cuda/sparse_matrix_operations.hpp"
#include "viennacl/linalg/cuda/iterative_operations.hpp"

namespace viennacl { namespace linalg { namespace b200 {

inline ViennaCLBackend backend()
{
  static ViennaCLBackend h = NULL;
  if (!h && ViennaCLBackendCreateOnDevice(&h, -1, (void*)cudaStreamLegacy) != ViennaCLSuccess) throw viennacl::memory_exception("libvcl_b200: no B200 device");
  return h;
}
inline void check(ViennaCLStatus s)
{ if (s != ViennaCLSuccess) throw viennacl::memory_exception(ViennaCLBackendLastError(backend())); }

// cuda_arg() does not add start() (linalg/cuda/common.hpp; callers add traits::start, cuda/iterative_operations.hpp:1927)
inline double       *ptr(vector_base<double> & v)       { return viennacl::cuda_arg(v) + v.start(); }
inline const double *ptr(vector_base<double> const & v) { return viennacl::cuda_arg(v) + v.start(); }

template<unsigned int A> ViennaCLCUDADcsr csr(compressed_matrix<double, A> const & M)
{ ViennaCLCUDADcsr a = { int(M.size1()), int(M.size2()), int(M.nnz()), viennacl::cuda_arg<unsigned int>(M.handle1()),
                         viennacl::cuda_arg<unsigned int>(M.handle2()), viennacl::cuda_arg<double>(M.handle()),
                         viennacl::cuda_arg<unsigned int>(M.handle3()), int(M.blocks1()) };   // the matrix's own row-block plan
  return a; }
template<typename IndexT> ViennaCLCUDADsell sell(sliced_ell_matrix<double, IndexT> const & M)
{ ViennaCLCUDADsell a = { int(M.size1()), int(M.size2()), int(M.rows_per_block()), viennacl::cuda_arg<unsigned int>(M.handle1()),
                          viennacl::cuda_arg<unsigned int>(M.handle2()), viennacl::cuda_arg<unsigned int>(M.handle3()),
                          viennacl::cuda_arg<double>(M.handle()), NULL };
  return a; }

// ---- linalg/sparse_matrix_operations.hpp: prod_impl -> cuda::prod_impl (cuda/sparse_matrix_operations.hpp:262-396, :2241-2289)
template<unsigned int A>
void prod_impl(compressed_matrix<double, A> const & M, vector_base<double> const & x, double alpha, vector_base<double> & y, double beta)
{ check(ViennaCLCUDADcsrmv(backend(), int(M.size1()), int(M.size2()), int(M.nnz()), viennacl::cuda_arg<unsigned int>(M.handle1()),
                           viennacl::cuda_arg<unsigned int>(M.handle2()), viennacl::cuda_arg<double>(M.handle()),
                           viennacl::cuda_arg<unsigned int>(M.handle3()), int(M.blocks1()),
                           viennacl::cuda_arg(x), int(x.start()), int(x.stride()), alpha, viennacl::cuda_arg(y), int(y.start()), int(y.stride()), beta)); }
template<typename IndexT> void prod_impl(sliced_ell_matrix<double, IndexT> const & M, vector_base<double> const & x, double alpha, vector_base<double> & y, double beta)
{ check(ViennaCLCUDADsellmv(backend(), int(M.size1()), int(M.size2()), int(M.rows_per_block()), viennacl::cuda_arg<unsigned int>(M.handle1()),
                            viennacl::cuda_arg<unsigned int>(M.handle2()), viennacl::cuda_arg<unsigned int>(M.handle3()), viennacl::cuda_arg<double>(M.handle()),
                            viennacl::cuda_arg(x), int(x.start()), int(x.stride()), alpha, viennacl::cuda_arg(y), int(y.start()), int(y.stride()), beta)); }
// any other matrix type / scalar type keeps the reference's CUDA kernels
template<typename MatrixT, typename NumericT>
void prod_impl(MatrixT const & M, vector_base<NumericT> const & x, NumericT alpha, vector_base<NumericT> & y, NumericT beta)
{ viennacl::linalg::cuda::prod_impl(M, x, alpha, y, beta); }

// ---- linalg/iterative_operations.hpp:77-79 -> cuda::pipelined_cg_vector_update (cuda/iterative_operations.hpp:43-103)
inline void pipelined_cg_vector_update(vector_base<double> & result, double alpha, vector_base<double> & p, vector_base<double> & r,
                                       vector_base<double> const & Ap, double beta, vector_base<double> & buf)
{ check(ViennaCLCUDADpipelined_cg_vector_update(backend(), int(result.size()), ptr(result), alpha, ptr(p), ptr(r), ptr(Ap), beta, ptr(buf), int(buf.size()))); }
// ---- :112-114 -> cuda::pipelined_cg_prod (cuda/iterative_operations.hpp:279-322, :609-629)
template<unsigned int A>
void pipelined_cg_prod(compressed_matrix<double, A> const & M, vector_base<double> const & p, vector_base<double> & Ap, vector_base<double> & buf)
{ ViennaCLCUDADcsr a = csr(M); check(ViennaCLCUDADpipelined_cg_prod_csr(backend(), &a, ptr(p), ptr(Ap), ptr(buf), int(buf.size()))); }
template<typename IndexT> void pipelined_cg_prod(sliced_ell_matrix<double, IndexT> const & M, vector_base<double> const & p, vector_base<double> & Ap, vector_base<double> & buf)
{ ViennaCLCUDADsell a = sell(M); check(ViennaCLCUDADpipelined_cg_prod_sell(backend(), &a, ptr(p), ptr(Ap), ptr(buf), int(buf.size()))); }
template<typename MatrixT, typename NumericT>
void pipelined_cg_prod(MatrixT const & M, vector_base<NumericT> const & p, vector_base<NumericT> & Ap, vector_base<NumericT> & buf)
{ viennacl::linalg::cuda::pipelined_cg_prod(M, p, Ap, buf); }

// ---- :151-153 -> cuda::pipelined_bicgstab_update_s (cuda/iterative_operations.hpp:733-808)
inline void pipelined_bicgstab_update_s(vector_base<double> & s, vector_base<double> & r, vector_base<double> const & Ap,
                                        vector_base<double> & buf, vcl_size_t chunk, vcl_size_t chunk_offset)
{ check(ViennaCLCUDADpipelined_bicgstab_update_s(backend(), int(s.size()), ptr(s), ptr(r), ptr(Ap), ptr(buf), int(chunk), int(chunk_offset))); }
// ---- :188-190 -> cuda::pipelined_bicgstab_vector_update (cuda/iterative_operations.hpp:810-886)
inline void pipelined_bicgstab_vector_update(vector_base<double> & result, double alpha, vector_base<double> & p, double omega,
                                             vector_base<double> const & s, vector_base<double> & residual, vector_base<double> const & As,
                                             double beta, vector_base<double> const & Ap, vector_base<double> const & r0star,
                                             vector_base<double> & buf, vcl_size_t chunk)
{ check(ViennaCLCUDADpipelined_bicgstab_vector_update(backend(), int(result.size()), ptr(result), alpha, ptr(p), omega, ptr(s), ptr(residual), ptr(As),
                                                      beta, ptr(Ap), ptr(r0star), ptr(buf), int(chunk))); }
// ---- :226-228 -> cuda::pipelined_bicgstab_prod (cuda/iterative_operations.hpp:888-1136)
template<unsigned int A>
void pipelined_bicgstab_prod(compressed_matrix<double, A> const & M, vector_base<double> const & p, vector_base<double> & Ap,
                             vector_base<double> const & r0star, vector_base<double> & buf, vcl_size_t chunk, vcl_size_t chunk_offset)
{ ViennaCLCUDADcsr a = csr(M); check(ViennaCLCUDADpipelined_bicgstab_prod_csr(backend(), &a, ptr(p), ptr(Ap), ptr(r0star), ptr(buf), int(chunk), int(chunk_offset))); }
template<typename IndexT> void pipelined_bicgstab_prod(sliced_ell_matrix<double, IndexT> const & M, vector_base<double> const & p, vector_base<double> & Ap,
                                    vector_base<double> const & r0star, vector_base<double> & buf, vcl_size_t chunk, vcl_size_t chunk_offset)
{ ViennaCLCUDADsell a = sell(M); check(ViennaCLCUDADpipelined_bicgstab_prod_sell(backend(), &a, ptr(p), ptr(Ap), ptr(r0star), ptr(buf), int(chunk), int(chunk_offset))); }
template<typename MatrixT, typename NumericT>
void pipelined_bicgstab_prod(MatrixT const & M, vector_base<NumericT> const & p, vector_base<NumericT> & Ap, vector_base<NumericT> const & r0star,
                             vector_base<NumericT> & buf, vcl_size_t chunk, vcl_size_t chunk_offset)
{ viennacl::linalg::cuda::pipelined_bicgstab_prod(M, p, Ap, r0star, buf, chunk, chunk_offset); }

// ---- :267-269 -> cuda::pipelined_gmres_normalize_vk (cuda/iterative_operations.hpp:1597-1688)
inline void pipelined_gmres_normalize_vk(vector_base<double> & v_k, vector_base<double> const & residual, vector_base<double> & R, vcl_size_t offset_in_R,
                                         vector_base<double> const & buf, vector_base<double> & r_dot_vk, vcl_size_t chunk, vcl_size_t chunk_offset)
{ check(ViennaCLCUDADpipelined_gmres_normalize_vk(backend(), int(v_k.size()), ptr(v_k), ptr(residual), ptr(R), int(offset_in_R), ptr(buf), ptr(r_dot_vk),
                                                  int(chunk), int(chunk_offset))); }
// ---- :303-305 -> cuda::pipelined_gmres_gram_schmidt_stage1 (cuda/iterative_operations.hpp:1690-1770)
inline void pipelined_gmres_gram_schmidt_stage1(vector_base<double> const & basis, vcl_size_t n, vcl_size_t internal_n, vcl_size_t k,
                                                vector_base<double> & vi_in_vk, vcl_size_t chunk)
{ check(ViennaCLCUDADpipelined_gmres_gram_schmidt_stage1(backend(), ptr(basis), int(n), int(internal_n), int(k), ptr(vi_in_vk), int(chunk))); }
// ---- :341-343 -> cuda::pipelined_gmres_gram_schmidt_stage2 (cuda/iterative_operations.hpp:1772-1860)
inline void pipelined_gmres_gram_schmidt_stage2(vector_base<double> & basis, vcl_size_t n, vcl_size_t internal_n, vcl_size_t k,
                                                vector_base<double> const & vi_in_vk, vector_base<double> & R, vcl_size_t krylov_dim,
                                                vector_base<double> & buf, vcl_size_t chunk)
{ check(ViennaCLCUDADpipelined_gmres_gram_schmidt_stage2(backend(), ptr(basis), int(n), int(internal_n), int(k), ptr(vi_in_vk), ptr(R), int(krylov_dim),
                                                         ptr(buf), int(chunk))); }
// ---- :374-376 -> cuda::pipelined_gmres_update_result (cuda/iterative_operations.hpp:1862-1925)
inline void pipelined_gmres_update_result(vector_base<double> & result, vector_base<double> const & residual, vector_base<double> const & basis,
                                          vcl_size_t n, vcl_size_t internal_n, vector_base<double> const & coefficients, vcl_size_t k)
{ check(ViennaCLCUDADpipelined_gmres_update_result(backend(), int(n), ptr(result), ptr(residual), ptr(basis), int(internal_n), ptr(coefficients), int(k))); }
// ---- :408-410 -> cuda::pipelined_gmres_prod (cuda/iterative_operations.hpp:1927-2045)
template<unsigned int A>
void pipelined_gmres_prod(compressed_matrix<double, A> const & M, vector_base<double> const & p, vector_base<double> & Ap, vector_base<double> & buf)
{ ViennaCLCUDADcsr a = csr(M); check(ViennaCLCUDADpipelined_gmres_prod_csr(backend(), &a, ptr(p), ptr(Ap), ptr(buf), int(buf.size()))); }
template<typename IndexT> void pipelined_gmres_prod(sliced_ell_matrix<double, IndexT> const & M, vector_base<double> const & p, vector_base<double> & Ap, vector_base<double> & buf)
{ ViennaCLCUDADsell a = sell(M); check(ViennaCLCUDADpipelined_gmres_prod_sell(backend(), &a, ptr(p), ptr(Ap), ptr(buf), int(buf.size()))); }
template<typename MatrixT, typename NumericT>
void pipelined_gmres_prod(MatrixT const & M, vector_base<NumericT> const & p, vector_base<NumericT> & Ap, vector_base<NumericT> & buf)
{ viennacl::linalg::cuda::pipelined_gmres_prod(M, p, Ap, buf); }

// float (and any type the library half is not bound for here) keeps the reference's CUDA kernels
template<typename NumericT>
void pipelined_cg_vector_update(vector_base<NumericT> & result, NumericT alpha, vector_base<NumericT> & p, vector_base<NumericT> & r,
                                vector_base<NumericT> const & Ap, NumericT beta, vector_base<NumericT> & buf)
{ viennacl::linalg::cuda::pipelined_cg_vector_update(result, alpha, p, r, Ap, beta, buf); }
template<typename NumericT>
void pipelined_bicgstab_update_s(vector_base<NumericT> & s, vector_base<NumericT> & r, vector_base<NumericT> const & Ap,
                                 vector_base<NumericT> & buf, vcl_size_t chunk, vcl_size_t chunk_offset)
{ viennacl::linalg::cuda::pipelined_bicgstab_update_s(s, r, Ap, buf, chunk, chunk_offset); }
template<typename NumericT>
void pipelined_bicgstab_vector_update(vector_base<NumericT> & result, NumericT alpha, vector_base<NumericT> & p, NumericT omega,
                                      vector_base<NumericT> const & s, vector_base<NumericT> & residual, vector_base<NumericT> const & As,
                                      NumericT beta, vector_base<NumericT> const & Ap, vector_base<NumericT> const & r0star,
                                      vector_base<NumericT> & buf, vcl_size_t chunk)
{ viennacl::linalg::cuda::pipelined_bicgstab_vector_update(result, alpha, p, omega, s, residual, As, beta, Ap, r0star, buf, chunk); }
template<typename NumericT>
void pipelined_gmres_normalize_vk(vector_base<NumericT> & v_k, vector_base<NumericT> const & residual, vector_base<NumericT> & R, vcl_size_t offset_in_R,
                                  vector_base<NumericT> const & buf, vector_base<NumericT> & r_dot_vk, vcl_size_t chunk, vcl_size_t chunk_offset)
{ viennacl::linalg::cuda::pipelined_gmres_normalize_vk(v_k, residual, R, offset_in_R, buf, r_dot_vk, chunk, chunk_offset); }
template<typename NumericT>
void pipelined_gmres_gram_schmidt_stage1(vector_base<NumericT> const & basis, vcl_size_t n, vcl_size_t internal_n, vcl_size_t k,
                                         vector_base<NumericT> & vi_in_vk, vcl_size_t chunk)
{ viennacl::linalg::cuda::pipelined_gmres_gram_schmidt_stage1(basis, n, internal_n, k, vi_in_vk, chunk); }
template<typename NumericT>
void pipelined_gmres_gram_schmidt_stage2(vector_base<NumericT> & basis, vcl_size_t n, vcl_size_t internal_n, vcl_size_t k,
                                         vector_base<NumericT> const & vi_in_vk, vector_base<NumericT> & R, vcl_size_t krylov_dim,
                                         vector_base<NumericT> & buf, vcl_size_t chunk)
{ viennacl::linalg::cuda::pipelined_gmres_gram_schmidt_stage2(basis, n, internal_n, k, vi_in_vk, R, krylov_dim, buf, chunk); }
template<typename NumericT>
void pipelined_gmres_update_result(vector_base<NumericT> & result, vector_base<NumericT> const & residual, vector_base<NumericT> const & basis,
                                   vcl_size_t n, vcl_size_t internal_n, vector_base<NumericT> const & coefficients, vcl_size_t k)
{ viennacl::linalg::cuda::pipelined_gmres_update_result(result, residual, basis, n, internal_n, coefficients, k); }

}}}
#endif
