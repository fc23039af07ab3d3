// viennacl/sliced_ell_matrix.hpp -- SELL-C-sigma (sigma = 1) matrix type (reference: sliced_ell_matrix.hpp:45-373).
// handle1() = u32 columns_per_block, handle2() = u32 column_indices, handle3() = u32 block_start, handle() = T elements;
// entry (r, j) at block_start[b] + j*C + (r mod C).  Conversion from CSR runs on the device (ViennaCLCUDADcsr2sell).
#ifndef VIENNACL_B200_SLICED_ELL_MATRIX_HPP
#define VIENNACL_B200_SLICED_ELL_MATRIX_HPP

#include "viennacl/forwards.h"
#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"

namespace viennacl
{

template<typename ScalarT, typename IndexT>
class sliced_ell_matrix
{
public:
  typedef backend::mem_handle handle_type;
  typedef ScalarT value_type;
  typedef vcl_size_t size_type;

  explicit sliced_ell_matrix() : rows_(0), cols_(0), rows_per_block_(0) {}
  sliced_ell_matrix(size_type num_rows, size_type num_cols, size_type num_rows_per_block_ = 0)
    : rows_(num_rows), cols_(num_cols), rows_per_block_(num_rows_per_block_) {}
  explicit sliced_ell_matrix(viennacl::context ctx) : rows_(0), cols_(0), rows_per_block_(0)
  { if (ctx.memory_type() != CUDA_MEMORY) throw memory_exception("only CUDA_MEMORY is available in the B200 build"); }

  vcl_size_t internal_size1() const { return tools::align_to_multiple<vcl_size_t>(rows_, rows_per_block_); }
  vcl_size_t internal_size2() const { return cols_; }
  vcl_size_t size1() const { return rows_; }
  vcl_size_t size2() const { return cols_; }
  vcl_size_t rows_per_block() const { return rows_per_block_; }

  handle_type & handle1() { return columns_per_block_; }
  const handle_type & handle1() const { return columns_per_block_; }
  handle_type & handle2() { return column_indices_; }
  const handle_type & handle2() const { return column_indices_; }
  handle_type & handle3() { return block_start_; }
  const handle_type & handle3() const { return block_start_; }
  handle_type & handle() { return elements_; }
  const handle_type & handle() const { return elements_; }

  void clear()
  {
    columns_per_block_ = handle_type(); column_indices_ = handle_type(); block_start_ = handle_type(); elements_ = handle_type();
  }

  typename viennacl::backend::b200::abi<ScalarT>::sell abi() const
  {
    typename viennacl::backend::b200::abi<ScalarT>::sell a = {ViennaCLInt(rows_), ViennaCLInt(cols_), ViennaCLInt(rows_per_block_), columns_per_block_.ptr<unsigned int>(),
                           column_indices_.ptr<unsigned int>(), block_start_.ptr<unsigned int>(), elements_.ptr<ScalarT>(),
                           sigma_ > 1 ? row_perm_.ptr<unsigned int>() : NULL};
    return a;
  }

  void vec_mul(vector_base<ScalarT> const & x, ScalarT alpha, vector_base<ScalarT> & y, ScalarT beta) const
  {
    assert(size1() == y.size() && size2() == x.size() && bool("Size check failed for sliced ELL matrix-vector product"));
    if (rows_ == 0) return;
    if (!columns_per_block_.get()) { detail::scale_by_beta(y, beta); return; }      // no entries (clear(), or never filled): A x = 0
    if (sigma_ > 1)
    {
      typename viennacl::backend::b200::abi<ScalarT>::sell a = abi();
      backend::b200::check(viennacl::backend::b200::abi<ScalarT>::sellmv_struct(backend::b200::handle(), &a, x.ptr(), ViennaCLInt(x.start()),
                                                                                 ViennaCLInt(x.stride()), alpha, y.ptr(), ViennaCLInt(y.start()),
                                                                                 ViennaCLInt(y.stride()), beta));
      return;
    }
    backend::b200::check(viennacl::backend::b200::abi<ScalarT>::sellmv(backend::b200::handle(), ViennaCLInt(rows_), ViennaCLInt(cols_), ViennaCLInt(rows_per_block_),
                                             columns_per_block_.ptr<unsigned int>(), column_indices_.ptr<unsigned int>(),
                                             block_start_.ptr<unsigned int>(), elements_.ptr<ScalarT>(),
                                             x.ptr(), ViennaCLInt(x.start()), ViennaCLInt(x.stride()), alpha,
                                             y.ptr(), ViennaCLInt(y.start()), ViennaCLInt(y.stride()), beta));
  }

  /** @brief Device-side conversion from CSR (layout of sliced_ell_matrix.hpp:140-214) */
  template<unsigned int AlignmentV>
  void from_csr(compressed_matrix<ScalarT, AlignmentV> const & A)
  {
    if (rows_per_block_ == 0) rows_per_block_ = 32;      // sliced_ell_matrix.hpp:146-147
    rows_ = A.size1(); cols_ = A.size2();
    if (rows_ == 0) return;
    vcl_size_t slices = (rows_ - 1) / rows_per_block_ + 1;
    columns_per_block_.create(sizeof(unsigned int) * slices);
    block_start_.create(sizeof(unsigned int) * slices);
    long long padded = 0;
    ViennaCLBackend b = backend::b200::handle();
    if (sigma_ > 1)
    {
      // SELL-C-sigma: rows sorted by length inside windows of sigma rows before slicing (extension; the reference has sigma = 1)
      row_perm_.create(sizeof(unsigned int) * slices * rows_per_block_);
      typedef viennacl::backend::b200::abi<ScalarT> abi_t;
      backend::b200::check(abi_t::csr2sell_sigma(b, ViennaCLInt(rows_), ViennaCLInt(rows_per_block_), ViennaCLInt(sigma_), A.handle1().template ptr<unsigned int>(),
                                                 A.handle2().template ptr<unsigned int>(), A.handle().template ptr<ScalarT>(), row_perm_.ptr<unsigned int>(),
                                                 columns_per_block_.ptr<unsigned int>(), block_start_.ptr<unsigned int>(), &padded, NULL, NULL));
      column_indices_.create(sizeof(unsigned int) * vcl_size_t(padded ? padded : 1));
      elements_.create(sizeof(ScalarT) * vcl_size_t(padded ? padded : 1));
      backend::b200::check(abi_t::csr2sell_sigma(b, ViennaCLInt(rows_), ViennaCLInt(rows_per_block_), ViennaCLInt(sigma_), A.handle1().template ptr<unsigned int>(),
                                                 A.handle2().template ptr<unsigned int>(), A.handle().template ptr<ScalarT>(), row_perm_.ptr<unsigned int>(),
                                                 columns_per_block_.ptr<unsigned int>(), block_start_.ptr<unsigned int>(), &padded,
                                                 column_indices_.ptr<unsigned int>(), elements_.ptr<ScalarT>()));
      padded_nnz_ = vcl_size_t(padded);
      return;
    }
    backend::b200::check(viennacl::backend::b200::abi<ScalarT>::csr2sell(b, ViennaCLInt(rows_), ViennaCLInt(rows_per_block_), A.handle1().template ptr<unsigned int>(),
                                               A.handle2().template ptr<unsigned int>(), A.handle().template ptr<ScalarT>(),
                                               columns_per_block_.ptr<unsigned int>(), block_start_.ptr<unsigned int>(), &padded, NULL, NULL));
    column_indices_.create(sizeof(unsigned int) * vcl_size_t(padded ? padded : 1));
    elements_.create(sizeof(ScalarT) * vcl_size_t(padded ? padded : 1));
    backend::b200::check(viennacl::backend::b200::abi<ScalarT>::csr2sell(b, ViennaCLInt(rows_), ViennaCLInt(rows_per_block_), A.handle1().template ptr<unsigned int>(),
                                               A.handle2().template ptr<unsigned int>(), A.handle().template ptr<ScalarT>(),
                                               columns_per_block_.ptr<unsigned int>(), block_start_.ptr<unsigned int>(), &padded,
                                               column_indices_.ptr<unsigned int>(), elements_.ptr<ScalarT>()));
  }

  /** @brief Sorting window of SELL-C-sigma for the NEXT copy() / from_csr(): a multiple of rows_per_block(), <= 4096; 1 (default)
   *  keeps the reference's layout.  Not in the reference, whose sigma is fixed at 1 (sliced_ell_matrix.hpp:43). */
  void sigma(vcl_size_t s) { sigma_ = s ? s : 1; }
  vcl_size_t sigma() const { return sigma_; }
  const handle_type & row_permutation() const { return row_perm_; }
  /** @brief Stored entries including padding (sigma > 1 only; 0 otherwise) */
  vcl_size_t padded_nnz() const { return padded_nnz_; }

private:
  vcl_size_t rows_, cols_, rows_per_block_;
  vcl_size_t sigma_ = 1, padded_nnz_ = 0;
  handle_type columns_per_block_, column_indices_, block_start_, elements_, row_perm_;
};

/** @brief Host (vector of maps) -> device SELL (sliced_ell_matrix.hpp:222-235): staged through a device CSR */
template<typename IndexT, typename NumericT, typename IndexT2>
void copy(std::vector< std::map<IndexT, NumericT> > const & cpu_matrix, sliced_ell_matrix<NumericT, IndexT2> & gpu_matrix)
{
  compressed_matrix<NumericT> csr;
  viennacl::copy(cpu_matrix, csr);
  gpu_matrix.from_csr(csr);
}

/** @brief Device CSR -> device SELL (extension; the reference only converts from host matrices) */
template<typename NumericT, unsigned int AlignmentV, typename IndexT2>
void copy(compressed_matrix<NumericT, AlignmentV> const & csr, sliced_ell_matrix<NumericT, IndexT2> & gpu_matrix) { gpu_matrix.from_csr(csr); }

namespace linalg
{
  template<typename NumericT, typename IndexT>
  void prod_impl(sliced_ell_matrix<NumericT, IndexT> const & mat, vector_base<NumericT> const & vec, NumericT alpha,
                 vector_base<NumericT> & result, NumericT beta)
  { mat.vec_mul(vec, alpha, result, beta); }
}

namespace traits
{
  template<typename T, typename I> vcl_size_t size1(sliced_ell_matrix<T, I> const & m) { return m.size1(); }
  template<typename T, typename I> vcl_size_t size2(sliced_ell_matrix<T, I> const & m) { return m.size2(); }
}

} // namespace viennacl
#endif
