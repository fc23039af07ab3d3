// viennacl/linalg/row_scaling.hpp -- row-scaling preconditioner object (reference: row_scaling.hpp:40-53, 150-190): a diagonal
// preconditioner whose entries are the inf-/1-/2-norms of the matrix rows.  Passed to solve(A, b, cg_tag | bicgstab_tag |
// gmres_tag, precond) on a compressed_matrix it selects the fused diagonal-preconditioner paths of the backend (the divide is
// folded into the solver kernels); apply() keeps the reference's stand-alone semantics.
#ifndef VIENNACL_B200_LINALG_ROW_SCALING_HPP
#define VIENNACL_B200_LINALG_ROW_SCALING_HPP
#include <stdexcept>
#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
namespace viennacl
{
namespace linalg
{
  /** @brief Tag holding the index p of the row norm: 0 sup-norm, 1 sum of absolute values, 2 Euclidean (default), :40-53 */
  class row_scaling_tag
  {
  public:
    row_scaling_tag(unsigned int p = 2) : norm_(p) {}
    unsigned int norm() const { return norm_; }
  private:
    unsigned int norm_;
  };

  class unknown_norm_exception : public std::runtime_error
  {
  public:
    explicit unknown_norm_exception(std::string const & what_arg) : std::runtime_error(what_arg) {}
  };

  template<typename MatrixT> class row_scaling;

  template<typename NumericT, unsigned int AlignmentV>
  class row_scaling< compressed_matrix<NumericT, AlignmentV> >
  {
  public:
    typedef compressed_matrix<NumericT, AlignmentV> matrix_type;
    row_scaling(matrix_type const & mat, row_scaling_tag const & tag) : diag_M_(mat.size1()), norm_(tag.norm()) { init(mat); }
    void init(matrix_type const & mat)
    {
      switch (norm_)                                       // row_scaling.hpp:168-181
      {
      case 0: detail::row_info(mat, diag_M_, detail::SPARSE_ROW_NORM_INF); break;
      case 1: detail::row_info(mat, diag_M_, detail::SPARSE_ROW_NORM_1); break;
      case 2: detail::row_info(mat, diag_M_, detail::SPARSE_ROW_NORM_2); break;
      default: throw unknown_norm_exception("Unknown norm when initializing row_scaling preconditioner!");
      }
    }
    template<unsigned int A2> void apply(viennacl::vector<NumericT, A2> & vec) const
    {
      assert(diag_M_.size() == vec.size() && bool("Size mismatch"));
      vec = element_div(vec, diag_M_);
    }
    unsigned int norm() const { return norm_; }
    /** @brief The backend's id of this preconditioner (ViennaCLB200PrecondRowScalingInf / 1 / 2) */
    ViennaCLB200Precond abi_id() const
    { return norm_ == 0 ? ViennaCLB200PrecondRowScalingInf : (norm_ == 1 ? ViennaCLB200PrecondRowScaling1 : ViennaCLB200PrecondRowScaling2); }
  private:
    viennacl::vector<NumericT> diag_M_;
    unsigned int norm_;
  };
}
}
#endif
