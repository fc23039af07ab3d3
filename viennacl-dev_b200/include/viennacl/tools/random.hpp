// viennacl/tools/random.hpp -- uniform_random_numbers<T> (reference: tools/random.hpp:44-70): values in [0, 1).
#ifndef VIENNACL_B200_TOOLS_RANDOM_HPP
#define VIENNACL_B200_TOOLS_RANDOM_HPP
#include <cstdlib>
namespace viennacl
{
namespace tools
{
  template<typename NumericT>
  class uniform_random_numbers
  {
  public:
    NumericT operator()() const { return static_cast<NumericT>(double(std::rand()) / (double(RAND_MAX) + 1.0)); }
  };
}
}
#endif
