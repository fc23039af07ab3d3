// viennacl/tools/adapter.hpp -- sparse_matrix_adapter / const_sparse_matrix_adapter (reference: tools/adapter.hpp:222-420):
// give a std::vector< std::map<IndexT, NumericT> > the size1() / size2() / operator()(i, j) interface of a matrix type, so
// that it can be filled like a uBLAS matrix and handed to viennacl::copy().
#ifndef VIENNACL_B200_TOOLS_ADAPTER_HPP
#define VIENNACL_B200_TOOLS_ADAPTER_HPP
#include <vector>
#include <map>
#include <cassert>
#include "viennacl/forwards.h"
namespace viennacl
{
namespace tools
{
  template<typename NumericT, typename SizeT = unsigned int>
  class const_sparse_matrix_adapter
  {
  public:
    typedef NumericT value_type;
    typedef std::vector< std::map<SizeT, NumericT> > container_type;
    const_sparse_matrix_adapter(container_type const & mat) : mat_(&mat), size1_(mat.size()), size2_(mat.size()) {}
    const_sparse_matrix_adapter(container_type const & mat, vcl_size_t num_rows, vcl_size_t num_cols) : mat_(&mat), size1_(num_rows), size2_(num_cols) {}
    vcl_size_t size1() const { return size1_; }
    vcl_size_t size2() const { return size2_; }
    NumericT operator()(vcl_size_t i, vcl_size_t j) const
    {
      typename std::map<SizeT, NumericT>::const_iterator it = (*mat_)[i].find(static_cast<SizeT>(j));
      return it == (*mat_)[i].end() ? NumericT(0) : it->second;
    }
    container_type const & get() const { return *mat_; }
  private:
    container_type const * mat_;
    vcl_size_t size1_, size2_;
  };

  template<typename NumericT, typename SizeT = unsigned int>
  class sparse_matrix_adapter : public const_sparse_matrix_adapter<NumericT, SizeT>
  {
    typedef const_sparse_matrix_adapter<NumericT, SizeT> base_type;
  public:
    typedef typename base_type::container_type container_type;
    sparse_matrix_adapter(container_type & mat) : base_type(mat), mat_(&mat), size1_(mat.size()), size2_(mat.size()) {}
    sparse_matrix_adapter(container_type & mat, vcl_size_t num_rows, vcl_size_t num_cols) : base_type(mat, num_rows, num_cols), mat_(&mat), size1_(num_rows), size2_(num_cols) {}
    NumericT & operator()(vcl_size_t i, vcl_size_t j) { return (*mat_)[i][static_cast<SizeT>(j)]; }
    vcl_size_t size1() const { return size1_; }
    vcl_size_t size2() const { return size2_; }
    void resize(vcl_size_t i, vcl_size_t j, bool preserve = true)
    {
      if (i > 0) mat_->resize(i); else mat_->clear();
      if (!preserve) for (vcl_size_t r = 0; r < mat_->size(); ++r) (*mat_)[r].clear();
      size1_ = i; size2_ = j;
    }
    void clear() { for (vcl_size_t r = 0; r < mat_->size(); ++r) (*mat_)[r].clear(); }
    container_type & get() { return *mat_; }
  private:
    container_type * mat_;
    vcl_size_t size1_, size2_;
  };
}
}
#endif
