// viennacl/backend/mem_handle.hpp -- ref-counted device buffer handle (reference: backend/mem_handle.hpp:89-245,
// backend/cuda.hpp:85-200) on top of the C-ABI allocator of libvcl_b200.so.
#ifndef VIENNACL_B200_BACKEND_MEM_HANDLE_HPP
#define VIENNACL_B200_BACKEND_MEM_HANDLE_HPP

#include <memory>
#include <string>
#include <cstdlib>
#include "viennacl/forwards.h"
#include "viennacl/context.hpp"
#include "vcl_b200.h"
#include "viennacl/backend/abi.hpp"

namespace viennacl
{
namespace backend
{
namespace cuda
{
  /** @brief Thrown for every non-zero status of the C-ABI (reference: backend/cuda.hpp:59-83 cuda_exception) */
  class cuda_exception : public std::runtime_error
  {
  public:
    cuda_exception(std::string const & what_arg, int err_code) : std::runtime_error(what_arg), error_code_(err_code) {}
    int error_code() const { return error_code_; }
  private:
    int error_code_;
  };
}

namespace b200
{
  /** @brief The process-wide backend handle (device = VCL_B200_DEVICE or the current device, own stream).
   *  The reference uses the implicit current device + default stream (SURVEY 8b "Threading"); user code that needs several
   *  devices or streams creates further handles through the C-ABI and installs one with set_handle(). */
  inline ViennaCLBackend & handle_slot() { static ViennaCLBackend h = NULL; return h; }

  inline ViennaCLBackend handle()
  {
    ViennaCLBackend & h = handle_slot();
    if (!h)
    {
      const char *dev = std::getenv("VCL_B200_DEVICE");
      ViennaCLStatus st = ViennaCLBackendCreateOnDevice(&h, dev ? std::atoi(dev) : -1, NULL);
      if (st != ViennaCLSuccess)
        throw cuda::cuda_exception("ViennaCL (B200 backend): no usable sm_100 device; this build has no CPU fallback", int(st));
    }
    return h;
  }

  inline void set_handle(ViennaCLBackend h) { handle_slot() = h; }

  inline void check(ViennaCLStatus st)
  {
    if (st == ViennaCLSuccess) return;
    std::string msg = std::string("ViennaCL (B200 backend) error: ") + ViennaCLBackendLastError(handle_slot());
    if (st == ViennaCLB200InvalidArgument) throw memory_exception(msg);
    throw cuda::cuda_exception(msg, int(st));
  }
}

/** @brief backend::finish(): blocks until all enqueued work is done (backend/memory.hpp:54-62) */
inline void finish() { b200::check(ViennaCLBackendSynchronize(b200::handle())); }

inline memory_types default_memory_type() { return CUDA_MEMORY; }

class mem_handle
{
  struct releaser
  {
    bool owned;
    void operator()(char *p) const { if (owned && p && b200::handle_slot()) ViennaCLCUDAMemFree(b200::handle_slot(), p); }
  };
public:
  typedef std::shared_ptr<char> cuda_handle_type;

  mem_handle() : active_handle_(MEMORY_NOT_INITIALIZED), size_in_bytes_(0) {}

  memory_types get_active_handle_id() const { return active_handle_; }
  void switch_active_handle_id(memory_types new_id)
  {
    if (new_id == MAIN_MEMORY || new_id == OPENCL_MEMORY)
      throw memory_exception("only CUDA_MEMORY is available in the B200 build (no host/OpenCL backend)");
    active_handle_ = new_id;
  }

  /** @brief Allocates `bytes` (optionally initialised from host memory): memory_create, backend/memory.hpp:87-130 */
  void create(vcl_size_t bytes, const void *host_ptr = NULL)
  {
    void *p = NULL;
    b200::check(ViennaCLCUDAMemAlloc(b200::handle(), &p, bytes ? bytes : 1));
    cuda_handle_ = cuda_handle_type(static_cast<char*>(p), releaser{true});
    active_handle_ = CUDA_MEMORY;
    size_in_bytes_ = bytes;
    if (host_ptr && bytes) b200::check(ViennaCLCUDAMemWrite(b200::handle(), p, 0, host_ptr, bytes, 0));
  }

  /** @brief Wraps user-provided device memory without taking ownership (the inc() idiom, vector.hpp:278-279) */
  void wrap(void *device_ptr, vcl_size_t bytes)
  {
    cuda_handle_ = cuda_handle_type(static_cast<char*>(device_ptr), releaser{false});
    active_handle_ = CUDA_MEMORY;
    size_in_bytes_ = bytes;
  }

  char *get() const { return cuda_handle_.get(); }
  template<typename T> T *ptr() const { return reinterpret_cast<T*>(cuda_handle_.get()); }
  cuda_handle_type & cuda_handle() { return cuda_handle_; }
  cuda_handle_type const & cuda_handle() const { return cuda_handle_; }

  vcl_size_t raw_size() const { return size_in_bytes_; }
  void raw_size(vcl_size_t new_size) { size_in_bytes_ = new_size; }

  bool operator==(mem_handle const & other) const { return cuda_handle_.get() == other.cuda_handle_.get(); }
  bool operator!=(mem_handle const & other) const { return !(*this == other); }
  bool operator<(mem_handle const & other) const { return cuda_handle_.get() < other.cuda_handle_.get(); }

  void swap(mem_handle & other)
  {
    std::swap(active_handle_, other.active_handle_);
    cuda_handle_.swap(other.cuda_handle_);
    std::swap(size_in_bytes_, other.size_in_bytes_);
  }

private:
  memory_types active_handle_;
  cuda_handle_type cuda_handle_;
  vcl_size_t size_in_bytes_;
};

// free-function spellings of the reference (backend/memory.hpp:87,140,177,220)
inline void memory_create(mem_handle & h, vcl_size_t size_in_bytes, viennacl::context const & = viennacl::context(), const void *host_ptr = NULL)
{ h.create(size_in_bytes, host_ptr); }
inline void memory_copy(mem_handle const & src, mem_handle & dst, vcl_size_t src_offset, vcl_size_t dst_offset, vcl_size_t bytes)
{ b200::check(ViennaCLCUDAMemCopy(b200::handle(), src.get(), src_offset, dst.get(), dst_offset, bytes)); }
inline void memory_write(mem_handle & dst, vcl_size_t dst_offset, vcl_size_t bytes, const void *ptr, bool async = false)
{ b200::check(ViennaCLCUDAMemWrite(b200::handle(), dst.get(), dst_offset, ptr, bytes, async ? 1 : 0)); }
inline void memory_read(mem_handle const & src, vcl_size_t src_offset, vcl_size_t bytes, void *ptr, bool async = false)
{ b200::check(ViennaCLCUDAMemRead(b200::handle(), src.get(), src_offset, ptr, bytes, async ? 1 : 0)); }

} // namespace backend

/** @brief Raw device pointer of an object's buffer: cuda_arg<T>() (linalg/cuda/common.hpp:39-146); does not add start() */
template<typename T> T *cuda_arg(backend::mem_handle const & h) { return h.ptr<T>(); }

} // namespace viennacl
#endif
