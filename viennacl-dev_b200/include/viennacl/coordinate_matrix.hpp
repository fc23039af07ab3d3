// viennacl/coordinate_matrix.hpp -- COO matrix type (reference: coordinate_matrix.hpp:186-400).  handle12() = u32 (row, col)
// pairs sorted by row, handle() = T elements -- the reference's layout.  Products and solvers run on a CSR index of the
// same entries (row pointers + column indices, values shared), built once on the device by ViennaCLCUDAcoo2csr; the
// reference's 64 group boundaries (handle3()) are an artefact of its segmented-reduction kernel and have no counterpart.
#ifndef VIENNACL_B200_COORDINATE_MATRIX_HPP
#define VIENNACL_B200_COORDINATE_MATRIX_HPP

#include <vector>
#include <map>
#include "viennacl/forwards.h"
#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"

namespace viennacl
{

template<typename NumericT, unsigned int AlignmentV>
class coordinate_matrix
{
public:
  typedef backend::mem_handle handle_type;
  typedef NumericT value_type;
  typedef vcl_size_t size_type;

  coordinate_matrix() : rows_(0), cols_(0), nonzeros_(0), row_block_num_(0) {}
  explicit coordinate_matrix(viennacl::context ctx) : rows_(0), cols_(0), nonzeros_(0), row_block_num_(0)
  { if (ctx.memory_type() != CUDA_MEMORY) throw memory_exception("only CUDA_MEMORY is available in the B200 build"); }
  coordinate_matrix(vcl_size_t rows, vcl_size_t cols, vcl_size_t nonzeros = 0, viennacl::context ctx = viennacl::context())
    : rows_(rows), cols_(cols), nonzeros_(nonzeros), row_block_num_(0)
  { if (ctx.memory_type() != CUDA_MEMORY) throw memory_exception("only CUDA_MEMORY is available in the B200 build"); }

  vcl_size_t size1() const { return rows_; }
  vcl_size_t size2() const { return cols_; }
  vcl_size_t nnz() const { return nonzeros_; }
  vcl_size_t internal_nnz() const { return nonzeros_; }

  const handle_type & handle12() const { return coord_buffer_; }
  const handle_type & handle() const { return elements_; }

  void clear()
  {
    nonzeros_ = 0; row_block_num_ = 0;
    coord_buffer_ = handle_type(); elements_ = handle_type(); idx_rows_ = handle_type(); idx_cols_ = handle_type(); idx_blocks_ = handle_type();
  }

  /** @brief Sets the matrix from host arrays: coords = (row, col) pairs sorted by row (coordinate_matrix.hpp:72-88) */
  void set(const unsigned int *coords, const NumericT *elements, vcl_size_t rows, vcl_size_t cols, vcl_size_t nonzeros)
  {
    rows_ = rows; cols_ = cols; nonzeros_ = nonzeros;
    coord_buffer_.create(sizeof(unsigned int) * 2 * (nonzeros ? nonzeros : 1), nonzeros ? coords : NULL);
    elements_.create(sizeof(NumericT) * (nonzeros ? nonzeros : 1), nonzeros ? elements : NULL);
    build_index();
  }

  /** @brief The CSR view products and solvers use (values shared with handle()) */
  typename viennacl::backend::b200::abi<NumericT>::csr abi() const
  {
    typename viennacl::backend::b200::abi<NumericT>::csr a = {ViennaCLInt(rows_), ViennaCLInt(cols_), ViennaCLInt(nonzeros_), idx_rows_.ptr<unsigned int>(),
                          idx_cols_.ptr<unsigned int>(), elements_.ptr<NumericT>(), idx_blocks_.ptr<unsigned int>(), ViennaCLInt(row_block_num_)};
    return a;
  }

  /** @brief y = alpha * A x + beta * y with the reference's COO arithmetic (host_based/sparse_matrix_operations.hpp:1222-1247) */
  void vec_mul(vector_base<NumericT> const & x, NumericT alpha, vector_base<NumericT> & y, NumericT beta) const
  {
    assert(size1() == y.size() && size2() == x.size() && bool("Size check failed for coordinate matrix-vector product"));
    if (rows_ == 0) return;
    if (nonzeros_ == 0 || !idx_rows_.get()) { detail::scale_by_beta(y, beta); return; }   // no entries (clear()): A x = 0
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::coomv(backend::b200::handle(), ViennaCLInt(rows_), ViennaCLInt(cols_), ViennaCLInt(nonzeros_),
                                            idx_rows_.ptr<unsigned int>(), idx_cols_.ptr<unsigned int>(), elements_.ptr<NumericT>(),
                                            idx_blocks_.ptr<unsigned int>(), ViennaCLInt(row_block_num_),
                                            x.ptr(), ViennaCLInt(x.start()), ViennaCLInt(x.stride()), alpha,
                                            y.ptr(), ViennaCLInt(y.start()), ViennaCLInt(y.stride()), beta));
  }

private:
  void build_index()
  {
    row_block_num_ = 0;
    if (rows_ == 0) return;
    ViennaCLBackend b = backend::b200::handle();
    idx_rows_.create(sizeof(unsigned int) * (rows_ + 1));
    idx_cols_.create(sizeof(unsigned int) * (nonzeros_ ? nonzeros_ : 1));
    backend::b200::check(ViennaCLCUDAcoo2csr(b, ViennaCLInt(rows_), ViennaCLInt(nonzeros_), coord_buffer_.ptr<unsigned int>(),
                                             idx_rows_.ptr<unsigned int>(), idx_cols_.ptr<unsigned int>()));
    ViennaCLInt nb = 0;
    backend::b200::check(ViennaCLCUDAcsr_row_blocks(b, ViennaCLInt(rows_), idx_rows_.ptr<unsigned int>(), NULL, &nb));
    idx_blocks_.create(sizeof(unsigned int) * (vcl_size_t(nb) + 1));
    backend::b200::check(ViennaCLCUDAcsr_row_blocks(b, ViennaCLInt(rows_), idx_rows_.ptr<unsigned int>(), idx_blocks_.ptr<unsigned int>(), &nb));
    row_block_num_ = vcl_size_t(nb);
  }

  vcl_size_t rows_, cols_, nonzeros_, row_block_num_;
  handle_type coord_buffer_, elements_, idx_rows_, idx_cols_, idx_blocks_;
};

/** @brief Host (vector of maps) -> device COO (coordinate_matrix.hpp:109-121); cols = max column + 1 unless the matrix was sized */
template<typename IndexT, typename NumericT, unsigned int AlignmentV>
void copy(std::vector< std::map<IndexT, NumericT> > const & cpu_matrix, coordinate_matrix<NumericT, AlignmentV> & gpu_matrix)
{
  vcl_size_t rows = cpu_matrix.size(), nnz = 0, max_col = 0;
  for (vcl_size_t i = 0; i < rows; ++i)
  {
    nnz += cpu_matrix[i].size();
    if (!cpu_matrix[i].empty()) max_col = std::max<vcl_size_t>(max_col, cpu_matrix[i].rbegin()->first);
  }
  if (rows == 0) return;
  std::vector<unsigned int> coords(2 * (nnz ? nnz : 1));
  std::vector<NumericT> va(nnz ? nnz : 1);
  vcl_size_t k = 0;
  for (vcl_size_t i = 0; i < rows; ++i)
    for (typename std::map<IndexT, NumericT>::const_iterator it = cpu_matrix[i].begin(); it != cpu_matrix[i].end(); ++it, ++k)
    { coords[2 * k] = static_cast<unsigned int>(i); coords[2 * k + 1] = static_cast<unsigned int>(it->first); va[k] = it->second; }
  vcl_size_t cols = gpu_matrix.size2() > 0 ? gpu_matrix.size2() : max_col + 1;
  gpu_matrix.set(&coords[0], &va[0], rows, cols, nnz);
}

/** @brief Device COO -> host vector of maps (coordinate_matrix.hpp:130-170) */
template<typename NumericT, unsigned int AlignmentV, typename IndexT>
void copy(coordinate_matrix<NumericT, AlignmentV> const & gpu_matrix, std::vector< std::map<IndexT, NumericT> > & cpu_matrix)
{
  cpu_matrix.assign(gpu_matrix.size1(), std::map<IndexT, NumericT>());
  const vcl_size_t nnz = gpu_matrix.nnz();
  if (nnz == 0) return;
  std::vector<unsigned int> coords(2 * nnz);
  std::vector<NumericT> va(nnz);
  backend::memory_read(gpu_matrix.handle12(), 0, sizeof(unsigned int) * 2 * nnz, &coords[0]);
  backend::memory_read(gpu_matrix.handle(), 0, sizeof(NumericT) * nnz, &va[0]);
  for (vcl_size_t k = 0; k < nnz; ++k) cpu_matrix[coords[2 * k]][static_cast<IndexT>(coords[2 * k + 1])] = va[k];
}

namespace linalg
{
  template<typename NumericT, unsigned int AlignmentV>
  void prod_impl(coordinate_matrix<NumericT, AlignmentV> const & mat, vector_base<NumericT> const & vec, NumericT alpha,
                 vector_base<NumericT> & result, NumericT beta)
  { mat.vec_mul(vec, alpha, result, beta); }
}

namespace traits
{
  template<typename T, unsigned int A> vcl_size_t size1(coordinate_matrix<T, A> const & m) { return m.size1(); }
  template<typename T, unsigned int A> vcl_size_t size2(coordinate_matrix<T, A> const & m) { return m.size2(); }
}

} // namespace viennacl
#endif
