// viennacl/linalg/prod.hpp -- lazy sparse matrix-vector product (reference: linalg/prod.hpp:350-361).
#ifndef VIENNACL_B200_LINALG_PROD_HPP
#define VIENNACL_B200_LINALG_PROD_HPP
#include "viennacl/forwards.h"
#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/sliced_ell_matrix.hpp"
#include "viennacl/ell_matrix.hpp"
#include "viennacl/hyb_matrix.hpp"
#include "viennacl/coordinate_matrix.hpp"
namespace viennacl
{
namespace linalg
{
  /** @brief y = prod(A, x), y += prod(A, x), y -= prod(A, x), x = prod(A, x): evaluated when assigned */
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::detail::matvec_expr<compressed_matrix<NumericT, AlignmentV>, NumericT>
  prod(compressed_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & x)
  { viennacl::detail::matvec_expr<compressed_matrix<NumericT, AlignmentV>, NumericT> e = {&A, &x}; return e; }

  template<typename NumericT, typename IndexT>
  viennacl::detail::matvec_expr<sliced_ell_matrix<NumericT, IndexT>, NumericT>
  prod(sliced_ell_matrix<NumericT, IndexT> const & A, vector_base<NumericT> const & x)
  { viennacl::detail::matvec_expr<sliced_ell_matrix<NumericT, IndexT>, NumericT> e = {&A, &x}; return e; }

  /** @brief User-defined (matrix-free) operators: any type with `apply(x, y)` and `size1()` (linalg/prod.hpp:350-361 + vector.hpp:3315-3319) */
  template<typename OperatorT, typename NumericT>
  viennacl::detail::matvec_expr<OperatorT, NumericT> prod(OperatorT const & A, vector_base<NumericT> const & x)
  { viennacl::detail::matvec_expr<OperatorT, NumericT> e = {&A, &x}; return e; }
}
}
#endif
