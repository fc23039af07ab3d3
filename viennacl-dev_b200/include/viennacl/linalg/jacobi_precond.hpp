// viennacl/linalg/jacobi_precond.hpp -- Jacobi preconditioner object (reference: jacobi_precond.hpp:43, 103-130).
// Passed to solve(A, b, bicgstab_tag, precond) it selects the fused left-preconditioned BiCGStab of the backend, where the
// divide by diag(A) is folded into the SpMV epilogue; apply() keeps the reference's stand-alone semantics.
#ifndef VIENNACL_B200_LINALG_JACOBI_PRECOND_HPP
#define VIENNACL_B200_LINALG_JACOBI_PRECOND_HPP
#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
namespace viennacl
{
namespace linalg
{
  class jacobi_tag {};

  template<typename MatrixT> class jacobi_precond;

  template<typename NumericT, unsigned int AlignmentV>
  class jacobi_precond< compressed_matrix<NumericT, AlignmentV> >
  {
  public:
    typedef compressed_matrix<NumericT, AlignmentV> matrix_type;
    jacobi_precond(matrix_type const & mat, jacobi_tag const &) : diag_A_(mat.size1()), mat_(&mat) { init(mat); }
    void init(matrix_type const & mat) { detail::row_info(mat, diag_A_, detail::SPARSE_ROW_DIAGONAL); mat_ = &mat; }
    template<unsigned int A2> void apply(viennacl::vector<NumericT, A2> & vec) const
    {
      assert(diag_A_.size() == vec.size() && bool("Size mismatch"));
      vec = element_div(vec, diag_A_);
    }
    viennacl::vector<NumericT> const & diagonal() const { return diag_A_; }
    matrix_type const * matrix() const { return mat_; }
  private:
    viennacl::vector<NumericT> diag_A_;
    matrix_type const * mat_;
  };
}
}
#endif
