// viennacl/ell_matrix.hpp -- ELLPACK matrix type (reference: ell_matrix.hpp:36-305).  handle2() = u32 coords, handle() = T
// elements, both internal_size1() * internal_maxnnz() entries, entry j of row r at j*internal_size1() + r; padding has
// value 0 and column 0.  Conversion from CSR runs on the device (ViennaCLCUDADcsr2ell).  AlignmentV = 1 only.
#ifndef VIENNACL_B200_ELL_MATRIX_HPP
#define VIENNACL_B200_ELL_MATRIX_HPP

#include "viennacl/forwards.h"
#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"

namespace viennacl
{

template<typename NumericT, unsigned int AlignmentV>
class ell_matrix
{
public:
  typedef backend::mem_handle handle_type;
  typedef NumericT value_type;
  typedef vcl_size_t size_type;

  ell_matrix() : rows_(0), cols_(0), maxnnz_(0) {}
  explicit ell_matrix(viennacl::context ctx) : rows_(0), cols_(0), maxnnz_(0)
  { if (ctx.memory_type() != CUDA_MEMORY) throw memory_exception("only CUDA_MEMORY is available in the B200 build"); }

  vcl_size_t internal_size1() const { return rows_; }
  vcl_size_t internal_size2() const { return cols_; }
  vcl_size_t size1() const { return rows_; }
  vcl_size_t size2() const { return cols_; }
  vcl_size_t internal_maxnnz() const { return maxnnz_; }
  vcl_size_t maxnnz() const { return maxnnz_; }
  vcl_size_t nnz() const { return rows_ * maxnnz_; }
  vcl_size_t internal_nnz() const { return internal_size1() * internal_maxnnz(); }

  handle_type & handle() { return elements_; }
  const handle_type & handle() const { return elements_; }
  handle_type & handle2() { return coords_; }
  const handle_type & handle2() const { return coords_; }

  void clear() { maxnnz_ = 0; coords_ = handle_type(); elements_ = handle_type(); }

  typename viennacl::backend::b200::abi<NumericT>::ell abi() const
  {
    typename viennacl::backend::b200::abi<NumericT>::ell a = {ViennaCLInt(rows_), ViennaCLInt(cols_), ViennaCLInt(rows_), ViennaCLInt(maxnnz_),
                          coords_.ptr<unsigned int>(), elements_.ptr<NumericT>()};
    return a;
  }

  /** @brief y = alpha * A x + beta * y (linalg/sparse_matrix_operations.hpp:90-121 -> cuda/...:1747-1838) */
  void vec_mul(vector_base<NumericT> const & x, NumericT alpha, vector_base<NumericT> & y, NumericT beta) const
  {
    assert(size1() == y.size() && size2() == x.size() && bool("Size check failed for ELL matrix-vector product"));
    if (rows_ == 0) return;
    if (maxnnz_ == 0 || !elements_.get()) { detail::scale_by_beta(y, beta); return; }   // no entries (clear()): A x = 0
    typename viennacl::backend::b200::abi<NumericT>::ell a = abi();
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::ellmv(backend::b200::handle(), &a, x.ptr(), ViennaCLInt(x.start()), ViennaCLInt(x.stride()), alpha,
                                            y.ptr(), ViennaCLInt(y.start()), ViennaCLInt(y.stride()), beta));
  }

  /** @brief Device-side conversion from CSR (layout of ell_matrix.hpp:122-166) */
  template<unsigned int A2>
  void from_csr(compressed_matrix<NumericT, A2> const & A)
  {
    rows_ = A.size1(); cols_ = A.size2(); maxnnz_ = 0;
    if (rows_ == 0) return;
    ViennaCLBackend b = backend::b200::handle();
    ViennaCLInt w = 0;
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::csr2ell(b, ViennaCLInt(rows_), A.handle1().template ptr<unsigned int>(), A.handle2().template ptr<unsigned int>(),
                                              A.handle().template ptr<NumericT>(), &w, NULL, NULL));
    maxnnz_ = vcl_size_t(w);
    const vcl_size_t tot = (rows_ * maxnnz_ > 0) ? rows_ * maxnnz_ : 1;
    coords_.create(sizeof(unsigned int) * tot);
    elements_.create(sizeof(NumericT) * tot);
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::csr2ell(b, ViennaCLInt(rows_), A.handle1().template ptr<unsigned int>(), A.handle2().template ptr<unsigned int>(),
                                              A.handle().template ptr<NumericT>(), &w, coords_.ptr<unsigned int>(), elements_.ptr<NumericT>()));
  }

private:
  vcl_size_t rows_, cols_, maxnnz_;
  handle_type coords_, elements_;
};

/** @brief Host (vector of maps) -> device ELL (ell_matrix.hpp:174-186): staged through a device CSR */
template<typename IndexT, typename NumericT, unsigned int AlignmentV>
void copy(std::vector< std::map<IndexT, NumericT> > const & cpu_matrix, ell_matrix<NumericT, AlignmentV> & gpu_matrix)
{
  compressed_matrix<NumericT> csr;
  viennacl::copy(cpu_matrix, csr);
  gpu_matrix.from_csr(csr);
}

/** @brief Device CSR -> device ELL (extension; the reference only converts from host matrices) */
template<typename NumericT, unsigned int A1, unsigned int A2>
void copy(compressed_matrix<NumericT, A1> const & csr, ell_matrix<NumericT, A2> & gpu_matrix) { gpu_matrix.from_csr(csr); }

/** @brief Device ELL -> host vector of maps (ell_matrix.hpp:194-240): zero-valued slots are padding */
template<typename NumericT, unsigned int AlignmentV, typename IndexT>
void copy(ell_matrix<NumericT, AlignmentV> const & gpu_matrix, std::vector< std::map<IndexT, NumericT> > & cpu_matrix)
{
  cpu_matrix.assign(gpu_matrix.size1(), std::map<IndexT, NumericT>());
  const vcl_size_t tot = gpu_matrix.internal_nnz();
  if (tot == 0) return;
  std::vector<unsigned int> co(tot);
  std::vector<NumericT> el(tot);
  backend::memory_read(gpu_matrix.handle2(), 0, sizeof(unsigned int) * tot, &co[0]);
  backend::memory_read(gpu_matrix.handle(), 0, sizeof(NumericT) * tot, &el[0]);
  for (vcl_size_t r = 0; r < gpu_matrix.size1(); ++r)
    for (vcl_size_t j = 0; j < gpu_matrix.internal_maxnnz(); ++j)
    {
      const vcl_size_t off = j * gpu_matrix.internal_size1() + r;
      if (el[off] > 0 || el[off] < 0) cpu_matrix[r][static_cast<IndexT>(co[off])] = el[off];
    }
}

namespace linalg
{
  template<typename NumericT, unsigned int AlignmentV>
  void prod_impl(ell_matrix<NumericT, AlignmentV> const & mat, vector_base<NumericT> const & vec, NumericT alpha,
                 vector_base<NumericT> & result, NumericT beta)
  { mat.vec_mul(vec, alpha, result, beta); }
}

namespace traits
{
  template<typename T, unsigned int A> vcl_size_t size1(ell_matrix<T, A> const & m) { return m.size1(); }
  template<typename T, unsigned int A> vcl_size_t size2(ell_matrix<T, A> const & m) { return m.size2(); }
}

} // namespace viennacl
#endif
