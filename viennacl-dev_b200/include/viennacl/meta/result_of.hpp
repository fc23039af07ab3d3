// viennacl/meta/result_of.hpp -- the two metafunctions the solver drivers and user code use (reference: meta/result_of.hpp:
// value_type<T> = element type of a vector / matrix, cpu_value_type<T> = the host scalar type behind it).
#ifndef VIENNACL_B200_META_RESULT_OF_HPP
#define VIENNACL_B200_META_RESULT_OF_HPP
#include <vector>
#include <map>
#include "viennacl/forwards.h"
namespace viennacl
{
template<typename NumericT> class vector_base;
template<typename NumericT, unsigned int AlignmentV> class vector;
template<typename NumericT, unsigned int AlignmentV> class compressed_matrix;
template<typename NumericT, unsigned int AlignmentV> class coordinate_matrix;
template<typename NumericT, unsigned int AlignmentV> class ell_matrix;
template<typename NumericT, unsigned int AlignmentV> class hyb_matrix;
template<typename NumericT, typename IndexT> class sliced_ell_matrix;

namespace result_of
{
  template<typename T> struct value_type { typedef typename T::value_type type; };
  template<> struct value_type<float> { typedef float type; };
  template<> struct value_type<double> { typedef double type; };

  template<typename T> struct cpu_value_type { typedef typename cpu_value_type<typename T::value_type>::type type; };
  template<> struct cpu_value_type<float> { typedef float type; };
  template<> struct cpu_value_type<double> { typedef double type; };
  template<> struct cpu_value_type<int> { typedef int type; };
  template<> struct cpu_value_type<unsigned int> { typedef unsigned int type; };
  template<typename T> struct cpu_value_type< vector_base<T> > { typedef T type; };
  template<typename T, unsigned int A> struct cpu_value_type< vector<T, A> > { typedef T type; };
  template<typename T, unsigned int A> struct cpu_value_type< compressed_matrix<T, A> > { typedef T type; };
  template<typename T, unsigned int A> struct cpu_value_type< coordinate_matrix<T, A> > { typedef T type; };
  template<typename T, unsigned int A> struct cpu_value_type< ell_matrix<T, A> > { typedef T type; };
  template<typename T, unsigned int A> struct cpu_value_type< hyb_matrix<T, A> > { typedef T type; };
  template<typename T, typename I> struct cpu_value_type< sliced_ell_matrix<T, I> > { typedef T type; };
}
}
#endif
