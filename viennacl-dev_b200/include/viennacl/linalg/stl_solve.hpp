// viennacl/linalg/stl_solve.hpp -- the STL convenience overloads of solve() (reference: cg.hpp:353-375, bicgstab.hpp:547-570,
// gmres.hpp:687-710): system given as std::vector< std::map<IndexT, NumericT> > and std::vector<NumericT>; the data is copied
// to the device, solved there and copied back.  Included by cg.hpp / bicgstab.hpp / gmres.hpp.
#ifndef VIENNACL_B200_LINALG_STL_SOLVE_HPP
#define VIENNACL_B200_LINALG_STL_SOLVE_HPP
#include <vector>
#include <map>
#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
namespace viennacl
{
namespace linalg
{
template<typename IndexT, typename NumericT, typename TagT, typename PreconditionerT>
std::vector<NumericT> solve(std::vector< std::map<IndexT, NumericT> > const & A, std::vector<NumericT> const & rhs, TagT const & tag,
                            PreconditionerT const & precond)
{
  viennacl::compressed_matrix<NumericT> vcl_A;
  viennacl::copy(A, vcl_A);
  viennacl::vector<NumericT> vcl_rhs(rhs.size());
  viennacl::copy(rhs, vcl_rhs);
  // unqualified: the overload for TagT is found by argument-dependent lookup at the point of instantiation, whichever of
  // cg.hpp / bicgstab.hpp / gmres.hpp was included first
  viennacl::vector<NumericT> vcl_result = solve(vcl_A, vcl_rhs, tag, precond);
  std::vector<NumericT> result(vcl_result.size());
  viennacl::copy(vcl_result, result);
  return result;
}

template<typename IndexT, typename NumericT, typename TagT>
std::vector<NumericT> solve(std::vector< std::map<IndexT, NumericT> > const & A, std::vector<NumericT> const & rhs, TagT const & tag)
{ return solve(A, rhs, tag, viennacl::linalg::no_precond()); }
}
}
#endif
