// viennacl/hyb_matrix.hpp -- hybrid ELL + CSR matrix type (reference: hyb_matrix.hpp:36-370).  handle() / handle2() = ELL
// elements / coords of width ell_nnz(), handle3() = u32 csr_rows[rows+1], handle4() = u32 csr_cols, handle5() = T
// csr_elements.  The ELL width is the smallest row length that covers at least csr_threshold() (default 0.8) of the rows
// (hyb_matrix.hpp:139-166); conversion from CSR runs on the device (ViennaCLCUDADcsr2hyb).  AlignmentV = 1 only.
#ifndef VIENNACL_B200_HYB_MATRIX_HPP
#define VIENNACL_B200_HYB_MATRIX_HPP

#include "viennacl/forwards.h"
#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"

namespace viennacl
{

template<typename NumericT, unsigned int AlignmentV>
class hyb_matrix
{
public:
  typedef backend::mem_handle handle_type;
  typedef NumericT value_type;
  typedef vcl_size_t size_type;

  hyb_matrix() : csr_threshold_(NumericT(0.8)), rows_(0), cols_(0), ellnnz_(0), csrnnz_(0) {}
  explicit hyb_matrix(viennacl::context ctx) : csr_threshold_(NumericT(0.8)), rows_(0), cols_(0), ellnnz_(0), csrnnz_(0)
  { if (ctx.memory_type() != CUDA_MEMORY) throw memory_exception("only CUDA_MEMORY is available in the B200 build"); }

  NumericT csr_threshold() const { return csr_threshold_; }
  void csr_threshold(NumericT thr) { csr_threshold_ = thr; }

  vcl_size_t internal_size1() const { return rows_; }
  vcl_size_t internal_size2() const { return cols_; }
  vcl_size_t size1() const { return rows_; }
  vcl_size_t size2() const { return cols_; }
  vcl_size_t internal_ellnnz() const { return ellnnz_; }
  vcl_size_t ell_nnz() const { return ellnnz_; }
  vcl_size_t csr_nnz() const { return csrnnz_; }

  const handle_type & handle() const { return ell_elements_; }
  const handle_type & handle2() const { return ell_coords_; }
  const handle_type & handle3() const { return csr_rows_; }
  const handle_type & handle4() const { return csr_cols_; }
  const handle_type & handle5() const { return csr_elements_; }

  void clear()
  {
    ellnnz_ = 0; csrnnz_ = 0;
    ell_coords_ = handle_type(); ell_elements_ = handle_type(); csr_rows_ = handle_type(); csr_cols_ = handle_type(); csr_elements_ = handle_type();
  }

  typename viennacl::backend::b200::abi<NumericT>::hyb abi() const
  {
    typename viennacl::backend::b200::abi<NumericT>::hyb a;
    a.ell.rows = ViennaCLInt(rows_); a.ell.cols = ViennaCLInt(cols_); a.ell.internal_rows = ViennaCLInt(rows_); a.ell.maxnnz = ViennaCLInt(ellnnz_);
    a.ell.coords = ell_coords_.ptr<unsigned int>(); a.ell.elements = ell_elements_.ptr<NumericT>();
    a.csr_rows = csr_rows_.ptr<unsigned int>(); a.csr_cols = csr_cols_.ptr<unsigned int>(); a.csr_elements = csr_elements_.ptr<NumericT>();
    a.csr_nnz = ViennaCLInt(csrnnz_);
    return a;
  }

  /** @brief y = alpha * A x + beta * y (linalg/sparse_matrix_operations.hpp:90-121 -> cuda/...:2298-2400) */
  void vec_mul(vector_base<NumericT> const & x, NumericT alpha, vector_base<NumericT> & y, NumericT beta) const
  {
    assert(size1() == y.size() && size2() == x.size() && bool("Size check failed for HYB matrix-vector product"));
    if (rows_ == 0) return;
    if (!ell_elements_.get() && !csr_rows_.get()) { detail::scale_by_beta(y, beta); return; }   // no entries (clear()): A x = 0
    typename viennacl::backend::b200::abi<NumericT>::hyb a = abi();
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::hybmv(backend::b200::handle(), &a, x.ptr(), ViennaCLInt(x.start()), ViennaCLInt(x.stride()), alpha,
                                            y.ptr(), ViennaCLInt(y.start()), ViennaCLInt(y.stride()), beta));
  }

  /** @brief Device-side conversion from CSR (layout and width rule of hyb_matrix.hpp:127-214) */
  template<unsigned int A2>
  void from_csr(compressed_matrix<NumericT, A2> const & A)
  {
    rows_ = A.size1(); cols_ = A.size2(); ellnnz_ = 0; csrnnz_ = 0;
    if (rows_ == 0) return;
    ViennaCLBackend b = backend::b200::handle();
    ViennaCLInt w = 0, tn = 0;
    const unsigned int *rp = A.handle1().template ptr<unsigned int>(), *ci = A.handle2().template ptr<unsigned int>();
    const NumericT *va = A.handle().template ptr<NumericT>();
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::csr2hyb(b, ViennaCLInt(rows_), ViennaCLInt(cols_), rp, ci, va, NumericT(csr_threshold_), &w, &tn,
                                              NULL, NULL, NULL, NULL, NULL));
    ellnnz_ = vcl_size_t(w); csrnnz_ = vcl_size_t(tn);
    const vcl_size_t tot = (rows_ * ellnnz_ > 0) ? rows_ * ellnnz_ : 1;
    ell_coords_.create(sizeof(unsigned int) * tot);
    ell_elements_.create(sizeof(NumericT) * tot);
    csr_rows_.create(sizeof(unsigned int) * (rows_ + 1));
    csr_cols_.create(sizeof(unsigned int) * (csrnnz_ ? csrnnz_ : 1));
    csr_elements_.create(sizeof(NumericT) * (csrnnz_ ? csrnnz_ : 1));
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::csr2hyb(b, ViennaCLInt(rows_), ViennaCLInt(cols_), rp, ci, va, NumericT(csr_threshold_), &w, &tn,
                                              ell_coords_.ptr<unsigned int>(), ell_elements_.ptr<NumericT>(), csr_rows_.ptr<unsigned int>(),
                                              csr_cols_.ptr<unsigned int>(), csr_elements_.ptr<NumericT>()));
  }

private:
  NumericT csr_threshold_;
  vcl_size_t rows_, cols_, ellnnz_, csrnnz_;
  handle_type ell_coords_, ell_elements_, csr_rows_, csr_cols_, csr_elements_;
};

/** @brief Host (vector of maps) -> device HYB (hyb_matrix.hpp:222-234): staged through a device CSR */
template<typename IndexT, typename NumericT, unsigned int AlignmentV>
void copy(std::vector< std::map<IndexT, NumericT> > const & cpu_matrix, hyb_matrix<NumericT, AlignmentV> & gpu_matrix)
{
  compressed_matrix<NumericT> csr;
  viennacl::copy(cpu_matrix, csr);
  gpu_matrix.from_csr(csr);
}

/** @brief Device CSR -> device HYB (extension; the reference only converts from host matrices) */
template<typename NumericT, unsigned int A1, unsigned int A2>
void copy(compressed_matrix<NumericT, A1> const & csr, hyb_matrix<NumericT, A2> & gpu_matrix) { gpu_matrix.from_csr(csr); }

namespace linalg
{
  template<typename NumericT, unsigned int AlignmentV>
  void prod_impl(hyb_matrix<NumericT, AlignmentV> const & mat, vector_base<NumericT> const & vec, NumericT alpha,
                 vector_base<NumericT> & result, NumericT beta)
  { mat.vec_mul(vec, alpha, result, beta); }
}

namespace traits
{
  template<typename T, unsigned int A> vcl_size_t size1(hyb_matrix<T, A> const & m) { return m.size1(); }
  template<typename T, unsigned int A> vcl_size_t size2(hyb_matrix<T, A> const & m) { return m.size2(); }
}

} // namespace viennacl
#endif
