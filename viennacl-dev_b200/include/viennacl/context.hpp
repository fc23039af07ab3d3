// viennacl/context.hpp -- memory domain tag carried by every object (context.hpp:39-83).  In this build the backend handle
// (device, stream, communicator) hangs off viennacl::backend::b200::handle(); context only records the memory type.
#ifndef VIENNACL_B200_CONTEXT_HPP
#define VIENNACL_B200_CONTEXT_HPP
#include "viennacl/forwards.h"
namespace viennacl
{
class context
{
public:
  context() : mem_type_(CUDA_MEMORY) {}
  explicit context(viennacl::memory_types mtype) : mem_type_(mtype)
  {
    if (mem_type_ == MEMORY_NOT_INITIALIZED) mem_type_ = CUDA_MEMORY;
  }
  viennacl::memory_types memory_type() const { return mem_type_; }
private:
  viennacl::memory_types mem_type_;
};
}
#endif
