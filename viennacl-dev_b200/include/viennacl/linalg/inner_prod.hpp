// viennacl/linalg/inner_prod.hpp -- <x, y> to a host scalar (reference: linalg/inner_prod.hpp, cuda/vector_operations.hpp:1551-1579).
#ifndef VIENNACL_B200_LINALG_INNER_PROD_HPP
#define VIENNACL_B200_LINALG_INNER_PROD_HPP
#include "viennacl/vector.hpp"
namespace viennacl
{
namespace linalg
{
  template<typename NumericT>
  viennacl::host_scalar<NumericT> inner_prod(vector_base<NumericT> const & x, vector_base<NumericT> const & y)
  {
    assert(x.size() == y.size() && bool("Incompatible vector sizes!"));
    NumericT r = 0;
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::dot(backend::b200::handle(), ViennaCLInt(x.size()), &r, x.ptr(), ViennaCLInt(x.start()), ViennaCLInt(x.stride()),
                                          y.ptr(), ViennaCLInt(y.start()), ViennaCLInt(y.stride())));
    return viennacl::host_scalar<NumericT>(r);
  }
}
}
#endif
