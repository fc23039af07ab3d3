// viennacl/linalg/norm_2.hpp -- ||x||_2 to a host scalar (reference: linalg/norm_2.hpp:108-119, cuda/vector_operations.hpp:2431-2448).
#ifndef VIENNACL_B200_LINALG_NORM_2_HPP
#define VIENNACL_B200_LINALG_NORM_2_HPP
#include "viennacl/vector.hpp"
namespace viennacl
{
namespace linalg
{
  template<typename NumericT>
  viennacl::host_scalar<NumericT> norm_2(vector_base<NumericT> const & x)
  {
    NumericT r = 0;
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::nrm2(backend::b200::handle(), ViennaCLInt(x.size()), &r, x.ptr(), ViennaCLInt(x.start()), ViennaCLInt(x.stride())));
    return viennacl::host_scalar<NumericT>(r);
  }
}
}
#endif
