// viennacl/compressed_matrix.hpp -- CSR matrix type of the B200 facade (reference: compressed_matrix.hpp:49-218, 628-1198,
// 1231-1318).  Array layout is the reference's: handle1() = u32 row_ptr[rows+1], handle2() = u32 col[nnz],
// handle() = T val[nnz], handle3() = u32 row_blocks[blocks1()+1].
#ifndef VIENNACL_B200_COMPRESSED_MATRIX_HPP
#define VIENNACL_B200_COMPRESSED_MATRIX_HPP

#include <vector>
#include <map>
#include "viennacl/forwards.h"
#include "viennacl/vector.hpp"
#include "viennacl/tools/adapter.hpp"

namespace viennacl
{

template<typename NumericT, unsigned int AlignmentV>
class compressed_matrix
{
public:
  typedef backend::mem_handle handle_type;
  typedef NumericT value_type;
  typedef vcl_size_t size_type;

  compressed_matrix() : rows_(0), cols_(0), nonzeros_(0), row_block_num_(0) {}
  explicit compressed_matrix(vcl_size_t rows, vcl_size_t cols, vcl_size_t nonzeros = 0, viennacl::context ctx = viennacl::context())
    : rows_(rows), cols_(cols), nonzeros_(nonzeros), row_block_num_(0)
  {
    check_ctx(ctx);
    if (rows_ > 0) { std::vector<unsigned int> z(rows_ + 1, 0u); row_buffer_.create(sizeof(unsigned int) * (rows_ + 1), &z[0]); }
    if (nonzeros_ > 0) { col_buffer_.create(sizeof(unsigned int) * nonzeros_); elements_.create(sizeof(NumericT) * nonzeros_); }
  }
  explicit compressed_matrix(vcl_size_t rows, vcl_size_t cols, viennacl::context ctx) : rows_(rows), cols_(cols), nonzeros_(0), row_block_num_(0) { check_ctx(ctx); }
  explicit compressed_matrix(viennacl::context ctx) : rows_(0), cols_(0), nonzeros_(0), row_block_num_(0) { check_ctx(ctx); }

  /** @brief Wraps existing CUDA buffers holding CSR arrays, no ownership taken (compressed_matrix.hpp:740-781) */
  explicit compressed_matrix(unsigned int *mem_row_buffer, unsigned int *mem_col_buffer, NumericT *mem_elements, viennacl::memory_types mem_type,
                             vcl_size_t rows, vcl_size_t cols, vcl_size_t nonzeros)
    : rows_(rows), cols_(cols), nonzeros_(nonzeros), row_block_num_(0)
  {
    if (mem_type != CUDA_MEMORY) throw memory_exception("only CUDA_MEMORY buffers can be wrapped in the B200 build");
    row_buffer_.wrap(mem_row_buffer, sizeof(unsigned int) * (rows + 1));
    col_buffer_.wrap(mem_col_buffer, sizeof(unsigned int) * nonzeros);
    elements_.wrap(mem_elements, sizeof(NumericT) * nonzeros);
    generate_row_block_information();
  }

  /** @brief Sets the CSR arrays from host memory (compressed_matrix.hpp:917-944) */
  void set(const void *row_jumper, const void *col_buffer, const NumericT *elements, vcl_size_t rows, vcl_size_t cols, vcl_size_t nonzeros)
  {
    assert((cols > 0) && (rows > 0) && bool("Invalid matrix dimensions"));
    rows_ = rows; cols_ = cols; nonzeros_ = nonzeros;
    row_buffer_.create(sizeof(unsigned int) * (rows + 1), row_jumper);
    col_buffer_.create(sizeof(unsigned int) * (nonzeros ? nonzeros : 1), nonzeros ? col_buffer : NULL);
    elements_.create(sizeof(NumericT) * (nonzeros ? nonzeros : 1), nonzeros ? elements : NULL);
    generate_row_block_information();
  }

  /** @brief Row blocks for the streaming SpMV kernel (compressed_matrix.hpp:1152-1188 analogue; built by the backend) */
  void generate_row_block_information()
  {
    row_block_num_ = 0;
    row_blocks_ = handle_type();
    if (rows_ == 0) return;
    ViennaCLInt nb = 0;
    backend::b200::check(ViennaCLCUDAcsr_row_blocks(backend::b200::handle(), ViennaCLInt(rows_), row_buffer_.ptr<unsigned int>(), NULL, &nb));
    row_blocks_.create(sizeof(unsigned int) * (vcl_size_t(nb) + 1));
    backend::b200::check(ViennaCLCUDAcsr_row_blocks(backend::b200::handle(), ViennaCLInt(rows_), row_buffer_.ptr<unsigned int>(),
                                                    row_blocks_.ptr<unsigned int>(), &nb));
    row_block_num_ = vcl_size_t(nb);
  }

  void clear()
  {
    nonzeros_ = 0; row_block_num_ = 0;
    if (rows_ > 0) { std::vector<unsigned int> z(rows_ + 1, 0u); row_buffer_.create(sizeof(unsigned int) * (rows_ + 1), &z[0]); }
    col_buffer_ = handle_type(); elements_ = handle_type(); row_blocks_ = handle_type();
    if (rows_ > 0) generate_row_block_information();
  }

  vcl_size_t size1() const { return rows_; }
  vcl_size_t size2() const { return cols_; }
  vcl_size_t nnz() const { return nonzeros_; }
  vcl_size_t blocks1() const { return row_block_num_; }

  const handle_type & handle1() const { return row_buffer_; }
  const handle_type & handle2() const { return col_buffer_; }
  const handle_type & handle3() const { return row_blocks_; }
  const handle_type & handle() const { return elements_; }
  handle_type & handle1() { return row_buffer_; }
  handle_type & handle2() { return col_buffer_; }
  handle_type & handle3() { return row_blocks_; }
  handle_type & handle() { return elements_; }

  viennacl::memory_types memory_context() const { return CUDA_MEMORY; }
  /** @brief compressed_matrix.hpp:1120-1141; this build has one memory domain, anything else is refused loudly */
  void switch_memory_context(viennacl::context new_ctx) { check_ctx(new_ctx); }

  /** @brief Read access to entry (i, j); 0 if it is not stored (compressed_matrix.hpp:1013-1040).  One small D2H per call. */
  NumericT operator()(vcl_size_t i, vcl_size_t j) const
  {
    assert(i < rows_ && j < cols_ && bool("index out of bounds"));
    unsigned int rp[2];
    backend::memory_read(row_buffer_, sizeof(unsigned int) * i, sizeof(unsigned int) * 2, rp);
    const vcl_size_t len = rp[1] - rp[0];
    if (len == 0) return NumericT(0);
    std::vector<unsigned int> ci(len);
    backend::memory_read(col_buffer_, sizeof(unsigned int) * rp[0], sizeof(unsigned int) * len, &ci[0]);
    for (vcl_size_t k = 0; k < len; ++k)
      if (ci[k] == j)
      {
        NumericT v;
        backend::memory_read(elements_, sizeof(NumericT) * (rp[0] + k), sizeof(NumericT), &v);
        return v;
      }
    return NumericT(0);
  }

  /** @brief Writes entry (i, j), inserting it if it is not stored yet (the entry_proxy assignment of compressed_matrix.hpp:
   *  1013-1100; an insertion rebuilds the arrays, exactly as costly as in the reference). */
  void set_entry(vcl_size_t i, vcl_size_t j, NumericT value)
  {
    assert(i < rows_ && j < cols_ && bool("index out of bounds"));
    std::vector<unsigned int> rp(rows_ + 1), ci(nonzeros_ ? nonzeros_ : 1);
    std::vector<NumericT> va(nonzeros_ ? nonzeros_ : 1);
    backend::memory_read(row_buffer_, 0, sizeof(unsigned int) * rp.size(), &rp[0]);
    if (nonzeros_ > 0)
    {
      backend::memory_read(col_buffer_, 0, sizeof(unsigned int) * nonzeros_, &ci[0]);
      backend::memory_read(elements_, 0, sizeof(NumericT) * nonzeros_, &va[0]);
    }
    unsigned int k = rp[i];
    while (k < rp[i + 1] && ci[k] < j) ++k;
    if (k < rp[i + 1] && ci[k] == j)
    {
      backend::memory_write(elements_, sizeof(NumericT) * k, sizeof(NumericT), &value);
      return;
    }
    ci.resize(nonzeros_); va.resize(nonzeros_);
    ci.insert(ci.begin() + k, static_cast<unsigned int>(j));
    va.insert(va.begin() + k, value);
    for (vcl_size_t r = i + 1; r <= rows_; ++r) rp[r] += 1;
    set(&rp[0], &ci[0], &va[0], rows_, cols_, nonzeros_ + 1);
  }

  /** @brief Resizes the matrix (compressed_matrix.hpp:946-1010): entries outside the new shape are dropped when `preserve`,
   *  otherwise the matrix is emptied. */
  void resize(vcl_size_t new_size1, vcl_size_t new_size2, bool preserve = true)
  {
    assert(new_size1 > 0 && new_size2 > 0 && bool("Cannot resize to zero size!"));
    std::vector<unsigned int> rp(rows_ + 1, 0u), ci(nonzeros_ ? nonzeros_ : 1);
    std::vector<NumericT> va(nonzeros_ ? nonzeros_ : 1);
    if (preserve && rows_ > 0)
    {
      backend::memory_read(row_buffer_, 0, sizeof(unsigned int) * rp.size(), &rp[0]);
      if (nonzeros_ > 0)
      {
        backend::memory_read(col_buffer_, 0, sizeof(unsigned int) * nonzeros_, &ci[0]);
        backend::memory_read(elements_, 0, sizeof(NumericT) * nonzeros_, &va[0]);
      }
    }
    std::vector<unsigned int> nrp(new_size1 + 1, 0u), nci;
    std::vector<NumericT> nva;
    for (vcl_size_t r = 0; r < new_size1; ++r)
    {
      if (preserve && r < rows_)
        for (unsigned int k = rp[r]; k < rp[r + 1]; ++k)
          if (ci[k] < new_size2) { nci.push_back(ci[k]); nva.push_back(va[k]); }
      nrp[r + 1] = static_cast<unsigned int>(nci.size());
    }
    const vcl_size_t nnz = nci.size();
    if (nci.empty()) { nci.push_back(0u); nva.push_back(NumericT(0)); }
    set(&nrp[0], &nci[0], &nva[0], new_size1, new_size2, nnz);
  }

  /** @brief The raw-array view the C-ABI takes */
  typename viennacl::backend::b200::abi<NumericT>::csr abi() const
  {
    typename viennacl::backend::b200::abi<NumericT>::csr a = {ViennaCLInt(rows_), ViennaCLInt(cols_), ViennaCLInt(nonzeros_), row_buffer_.ptr<unsigned int>(),
                          col_buffer_.ptr<unsigned int>(), elements_.ptr<NumericT>(), row_blocks_.ptr<unsigned int>(), ViennaCLInt(row_block_num_)};
    return a;
  }

  /** @brief y = alpha * A x + beta * y: linalg::prod_impl (linalg/sparse_matrix_operations.hpp:90-121) */
  void vec_mul(vector_base<NumericT> const & x, NumericT alpha, vector_base<NumericT> & y, NumericT beta) const
  {
    assert(size1() == y.size() && size2() == x.size() && bool("Size check failed for compressed matrix-vector product"));
    if (rows_ == 0) return;
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::csrmv(backend::b200::handle(), ViennaCLInt(rows_), ViennaCLInt(cols_), ViennaCLInt(nonzeros_),
                                            row_buffer_.ptr<unsigned int>(), col_buffer_.ptr<unsigned int>(), elements_.ptr<NumericT>(),
                                            row_blocks_.ptr<unsigned int>(), ViennaCLInt(row_block_num_),
                                            x.ptr(), ViennaCLInt(x.start()), ViennaCLInt(x.stride()), alpha,
                                            y.ptr(), ViennaCLInt(y.start()), ViennaCLInt(y.stride()), beta));
  }

private:
  void check_ctx(viennacl::context const & ctx) const
  { if (ctx.memory_type() != CUDA_MEMORY) throw memory_exception("only CUDA_MEMORY is available in the B200 build (no host/OpenCL backend)"); }

  vcl_size_t rows_, cols_, nonzeros_, row_block_num_;
  handle_type row_buffer_, row_blocks_, col_buffer_, elements_;
};

template<typename IndexT, typename NumericT, unsigned int AlignmentV>
void copy(std::vector< std::map<IndexT, NumericT> > const & cpu_matrix, compressed_matrix<NumericT, AlignmentV> & gpu_matrix);

/** @brief tools::(const_)sparse_matrix_adapter -> device CSR (the generic CPUMatrixT overload of compressed_matrix.hpp:49-105
 *  for the adapter of tools/adapter.hpp); the adapter's dimensions are kept */
template<typename NumericT, typename SizeT, unsigned int AlignmentV>
void copy(tools::const_sparse_matrix_adapter<NumericT, SizeT> const & cpu_matrix, compressed_matrix<NumericT, AlignmentV> & gpu_matrix)
{
  viennacl::copy(cpu_matrix.get(), gpu_matrix);
  if (gpu_matrix.size1() != cpu_matrix.size1() || gpu_matrix.size2() != cpu_matrix.size2())
    gpu_matrix.resize(cpu_matrix.size1(), cpu_matrix.size2(), true);
}
template<typename NumericT, typename SizeT, unsigned int AlignmentV>
void copy(tools::sparse_matrix_adapter<NumericT, SizeT> const & cpu_matrix, compressed_matrix<NumericT, AlignmentV> & gpu_matrix)
{ viennacl::copy(static_cast<tools::const_sparse_matrix_adapter<NumericT, SizeT> const &>(cpu_matrix), gpu_matrix); }

/** @brief Host (vector of maps) -> device CSR (compressed_matrix.hpp:190-218); cols = max column + 1 unless the matrix was sized */
template<typename IndexT, typename NumericT, unsigned int AlignmentV>
void copy(std::vector< std::map<IndexT, NumericT> > const & cpu_matrix, compressed_matrix<NumericT, AlignmentV> & gpu_matrix)
{
  vcl_size_t rows = cpu_matrix.size(), nnz = 0, max_col = 0;
  for (vcl_size_t i = 0; i < rows; ++i)
  {
    nnz += cpu_matrix[i].size();
    if (!cpu_matrix[i].empty()) max_col = std::max<vcl_size_t>(max_col, cpu_matrix[i].rbegin()->first);
  }
  if (rows == 0) return;
  std::vector<unsigned int> rp(rows + 1), ci(nnz ? nnz : 1);
  std::vector<NumericT> va(nnz ? nnz : 1);
  vcl_size_t k = 0;
  for (vcl_size_t i = 0; i < rows; ++i)
  {
    rp[i] = static_cast<unsigned int>(k);
    for (typename std::map<IndexT, NumericT>::const_iterator it = cpu_matrix[i].begin(); it != cpu_matrix[i].end(); ++it, ++k)
    { ci[k] = static_cast<unsigned int>(it->first); va[k] = it->second; }
  }
  rp[rows] = static_cast<unsigned int>(k);
  vcl_size_t cols = gpu_matrix.size2() > 0 ? gpu_matrix.size2() : max_col + 1;
  gpu_matrix.set(&rp[0], &ci[0], &va[0], rows, cols, nnz);
}

/** @brief Device CSR -> host vector of maps (compressed_matrix.hpp:406-460) */
template<typename NumericT, unsigned int AlignmentV, typename IndexT>
void copy(compressed_matrix<NumericT, AlignmentV> const & gpu_matrix, std::vector< std::map<IndexT, NumericT> > & cpu_matrix)
{
  cpu_matrix.assign(gpu_matrix.size1(), std::map<IndexT, NumericT>());
  if (gpu_matrix.size1() == 0) return;
  std::vector<unsigned int> rp(gpu_matrix.size1() + 1), ci(gpu_matrix.nnz());
  std::vector<NumericT> va(gpu_matrix.nnz());
  backend::memory_read(gpu_matrix.handle1(), 0, sizeof(unsigned int) * rp.size(), &rp[0]);
  if (gpu_matrix.nnz() > 0)
  {
    backend::memory_read(gpu_matrix.handle2(), 0, sizeof(unsigned int) * ci.size(), &ci[0]);
    backend::memory_read(gpu_matrix.handle(), 0, sizeof(NumericT) * va.size(), &va[0]);
  }
  for (vcl_size_t i = 0; i < gpu_matrix.size1(); ++i)
    for (unsigned int k = rp[i]; k < rp[i + 1]; ++k) cpu_matrix[i][static_cast<IndexT>(ci[k])] = va[k];
}

namespace linalg
{
  /** @brief result = alpha * A * vec + beta * result (linalg/sparse_matrix_operations.hpp:90-121) */
  template<typename NumericT, unsigned int AlignmentV>
  void prod_impl(compressed_matrix<NumericT, AlignmentV> const & mat, vector_base<NumericT> const & vec, NumericT alpha,
                 vector_base<NumericT> & result, NumericT beta)
  { mat.vec_mul(vec, alpha, result, beta); }

  namespace detail
  {
    enum row_info_types { SPARSE_ROW_NORM_INF = 0, SPARSE_ROW_NORM_1, SPARSE_ROW_NORM_2, SPARSE_ROW_DIAGONAL };

    /** @brief Per-row norms / diagonal (linalg/sparse_matrix_operations.hpp:48-74) */
    template<typename NumericT, unsigned int AlignmentV>
    void row_info(compressed_matrix<NumericT, AlignmentV> const & mat, vector_base<NumericT> & vec, row_info_types info_selector)
    {
      assert(vec.size() == mat.size1() && vec.stride() == 1 && bool("row_info needs a contiguous vector of size1() entries"));
      backend::b200::check(viennacl::backend::b200::abi<NumericT>::csr_row_info(backend::b200::handle(), ViennaCLInt(mat.size1()), mat.handle1().template ptr<unsigned int>(),
                                                     mat.handle2().template ptr<unsigned int>(), mat.handle().template ptr<NumericT>(),
                                                     vec.ptr() + vec.start(), ViennaCLInt(info_selector)));
    }
  }
}

namespace traits
{
  template<typename T, unsigned int A> vcl_size_t size1(compressed_matrix<T, A> const & m) { return m.size1(); }
  template<typename T, unsigned int A> vcl_size_t size2(compressed_matrix<T, A> const & m) { return m.size2(); }
}

} // namespace viennacl
#endif
