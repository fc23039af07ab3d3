// viennacl/forwards.h -- B200-native facade: names and enums the reference exposes to user code (forwards.h:369-374,
// :596-621, :897-902), nothing else.  Only double precision and CUDA_MEMORY are functional in this build: there is no host
// or OpenCL backend behind these headers.
#ifndef VIENNACL_B200_FORWARDS_H
#define VIENNACL_B200_FORWARDS_H

#include <cstddef>
#include <stdexcept>
#include <string>

#define VIENNACL_WITH_CUDA 1
#define VIENNACL_B200 1

namespace viennacl
{
typedef std::size_t vcl_size_t;
typedef std::ptrdiff_t vcl_ptrdiff_t;

enum memory_types { MEMORY_NOT_INITIALIZED, MAIN_MEMORY, OPENCL_MEMORY, CUDA_MEMORY };   // forwards.h:369-374

const vcl_size_t dense_padding_size = 128;                                               // forwards.h:385

class memory_exception : public std::exception                                           // forwards.h:596
{
public:
  memory_exception() : message_() {}
  memory_exception(std::string message) : message_("ViennaCL: Internal memory error: " + message) {}
  virtual const char* what() const throw() { return message_.c_str(); }
  virtual ~memory_exception() throw() {}
private:
  std::string message_;
};

class cuda_not_available_exception : public std::exception                               // forwards.h:609
{
public:
  cuda_not_available_exception() : message_("ViennaCL was compiled without CUDA support, but CUDA functionality required for this operation.") {}
  virtual const char* what() const throw() { return message_.c_str(); }
  virtual ~cuda_not_available_exception() throw() {}
private:
  std::string message_;
};

class zero_on_diagonal_exception : public std::runtime_error                             // forwards.h:621
{
public:
  zero_on_diagonal_exception(std::string const & what_arg) : std::runtime_error(what_arg) {}
};

template<typename NumericT> class vector_base;
template<typename NumericT, unsigned int AlignmentV = 1> class vector;
template<typename NumericT, unsigned int AlignmentV = 1> class compressed_matrix;
template<typename NumericT, typename IndexT = unsigned int> class sliced_ell_matrix;
template<typename NumericT, unsigned int AlignmentV = 1> class ell_matrix;
template<typename NumericT, unsigned int AlignmentV = 1> class hyb_matrix;
template<typename NumericT, unsigned int AlignmentV = 128> class coordinate_matrix;

namespace linalg
{
  /** @brief A tag class representing the use of no preconditioner (forwards.h:897-902) */
  class no_precond
  {
  public:
    template<typename VectorT> void apply(VectorT &) const {}
  };
}

} // namespace viennacl
#endif
