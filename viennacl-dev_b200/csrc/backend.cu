// backend.cu -- backend handle + device memory runtime of libvcl_b200.so.
// Replaces viennacl/backend/cuda.hpp:103-200 and the handle plumbing of backend/mem_handle.hpp:89-245 /
// backend/memory.hpp:54-367 for the CUDA domain (the C++ facade keeps mem_handle and calls these).
#include "common.cuh"
#include "prec.cuh"
#include "solver_state.cuh"
#include <cstdlib>
using vcl_f64::SolverState;       // the double-precision state is the larger one: both builds fit

ViennaCLStatus vcl_fail(ViennaCLBackend b, ViennaCLStatus st, const char *what, const char *file, int line)
{
  if (b)
  {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s (%s:%d)", what, file, line);
    b->last_error = buf;
  }
  return st;
}

ViennaCLStatus vcl_cuda_fail(ViennaCLBackend b, cudaError_t e, const char *what, const char *file, int line)
{
  if (b)
  {
    char buf[768];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) in %s (%s:%d)", (int)e, cudaGetErrorString(e), what, file, line);
    b->last_error = buf;
  }
  if (e == cudaErrorMemoryAllocation) return ViennaCLB200OutOfMemory;
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return ViennaCLB200NoDevice;
  return ViennaCLB200CudaError;
}

ViennaCLStatus vcl_ws_reserve(ViennaCLBackend b, size_t bytes)
{
  if (b->ws_bytes >= bytes) return ViennaCLSuccess;
  if (b->ws) { VCL_CUDA(b, cudaStreamSynchronize(b->stream)); VCL_CUDA(b, cudaFree(b->ws)); b->ws = nullptr; b->ws_bytes = 0; }
  VCL_CUDA(b, cudaMalloc(&b->ws, bytes));
  b->ws_bytes = bytes;
  return ViennaCLSuccess;
}

static ViennaCLStatus backend_init(ViennaCLBackend b, int device, void *stream)
{
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return vcl_fail(b, ViennaCLB200NoDevice, "no CUDA device available: libvcl_b200 has no CPU fallback", __FILE__, __LINE__);
  if (device < 0) { VCL_CUDA(b, cudaGetDevice(&device)); }
  VCL_REQUIRE(b, device < count, "device index out of range");
  VCL_CUDA(b, cudaSetDevice(device));
  cudaDeviceProp prop;
  VCL_CUDA(b, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return vcl_fail(b, ViennaCLB200NoDevice, "libvcl_b200 is built for sm_100a (B200) only", __FILE__, __LINE__);
  b->device = device;
  b->sm_count = prop.multiProcessorCount;
  b->l2_bytes = (size_t)prop.l2CacheSize;
  { int v = 0; b->coop_launch = (cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, device) == cudaSuccess && v) ? 1 : 0; }
  { const char *e = getenv("VCL_B200_PERSISTENT_ROWS"); if (e) b->persistent_rows = atoll(e); }
  if (stream) { b->stream = (cudaStream_t)stream; b->owns_stream = false; }
  else { VCL_CUDA(b, cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking)); b->owns_stream = true; }
  {
    int lo = 0, hi = 0;                       // communication outranks compute: its few CTAs must not queue behind SpMV CTAs
    VCL_CUDA(b, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    VCL_CUDA(b, cudaStreamCreateWithPriority(&b->comm_stream, cudaStreamNonBlocking, hi));
  }
  VCL_CUDA(b, cudaEventCreateWithFlags(&b->ev_a, cudaEventDisableTiming));
  VCL_CUDA(b, cudaEventCreateWithFlags(&b->ev_b, cudaEventDisableTiming));
  VCL_CUDA(b, cudaEventCreate(&b->tm_begin));
  VCL_CUDA(b, cudaEventCreate(&b->tm_end));
  VCL_CUDA(b, cudaMalloc(&b->partials, sizeof(double) * VCL_MAX_QUANT * VCL_MAX_BLOCKS));
  VCL_CUDA(b, cudaMalloc(&b->tickets, sizeof(unsigned int) * 16));
  VCL_CUDA(b, cudaMemset(b->tickets, 0, sizeof(unsigned int) * 16));
  VCL_CUDA(b, cudaMalloc(&b->dscal, sizeof(double) * VCL_DSCAL_COUNT));
  VCL_CUDA(b, cudaMemset(b->dscal, 0, sizeof(double) * VCL_DSCAL_COUNT));
  VCL_CUDA(b, cudaMallocHost(&b->hscal, sizeof(double) * VCL_DSCAL_COUNT));
  VCL_CUDA(b, cudaMalloc(&b->dstate, sizeof(SolverState)));
  VCL_CUDA(b, cudaMemset(b->dstate, 0, sizeof(SolverState)));
  VCL_CUDA(b, cudaMallocHost(&b->hstate, sizeof(SolverState)));
  VCL_CUDA(b, cudaDeviceSynchronize());
  return ViennaCLSuccess;
}

extern "C" {

const char *ViennaCLB200Version(void) { return "vcl_b200 0.1 (sm_100a)"; }

ViennaCLStatus ViennaCLBackendCreate(ViennaCLBackend *backend)
{
  return ViennaCLBackendCreateOnDevice(backend, -1, nullptr);
}

ViennaCLStatus ViennaCLBackendCreateOnDevice(ViennaCLBackend *backend, ViennaCLInt device, void *cuda_stream)
{
  if (!backend) return ViennaCLB200InvalidArgument;
  ViennaCLBackend b = new ViennaCLBackend_impl();
  ViennaCLStatus st = backend_init(b, device, cuda_stream);
  if (st != ViennaCLSuccess)
  {
    fprintf(stderr, "libvcl_b200: backend creation failed: %s\n", b->last_error.c_str());
    delete b;
    *backend = nullptr;
    return st;
  }
  *backend = b;
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLBackendDestroy(ViennaCLBackend *backend)
{
  if (!backend || !*backend) return ViennaCLSuccess;
  ViennaCLBackend b = *backend;
  cudaSetDevice(b->device);
  cudaStreamSynchronize(b->stream);
  ViennaCLBackendCommDestroy(b);
  if (b->ws) cudaFree(b->ws);
  if (b->flush_buf) cudaFree(b->flush_buf);
  cudaFree(b->partials); cudaFree(b->tickets); cudaFree(b->dscal); cudaFreeHost(b->hscal);
  cudaFree(b->dstate); cudaFreeHost(b->hstate);
  cudaEventDestroy(b->ev_a); cudaEventDestroy(b->ev_b); cudaEventDestroy(b->tm_begin); cudaEventDestroy(b->tm_end);
  cudaStreamDestroy(b->comm_stream);
  if (b->owns_stream) cudaStreamDestroy(b->stream);
  delete b;
  *backend = nullptr;
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLBackendSynchronize(ViennaCLBackend b)
{
  VCL_CHECK_BACKEND(b);
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLBackendGetStream(ViennaCLBackend b, void **s)
{
  VCL_CHECK_BACKEND(b);
  *s = (void*)b->stream;
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLBackendGetDevice(ViennaCLBackend b, ViennaCLInt *device, ViennaCLInt *sm_count)
{
  VCL_CHECK_BACKEND(b);
  if (device) *device = b->device;
  if (sm_count) *sm_count = b->sm_count;
  return ViennaCLSuccess;
}

const char *ViennaCLBackendLastError(ViennaCLBackend b) { return b ? b->last_error.c_str() : "backend not initialised"; }

ViennaCLStatus ViennaCLBackendTimerBegin(ViennaCLBackend b)
{
  VCL_CHECK_BACKEND(b);
  VCL_CUDA(b, cudaEventRecord(b->tm_begin, b->stream));
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLBackendTimerEnd(ViennaCLBackend b, double *ms)
{
  VCL_CHECK_BACKEND(b);
  VCL_CUDA(b, cudaEventRecord(b->tm_end, b->stream));
  VCL_CUDA(b, cudaEventSynchronize(b->tm_end));
  float f = 0;
  VCL_CUDA(b, cudaEventElapsedTime(&f, b->tm_begin, b->tm_end));
  if (ms) *ms = (double)f;
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLBackendFlushL2(ViennaCLBackend b)
{
  VCL_CHECK_BACKEND(b);
  if (!b->flush_buf)
  {
    b->flush_bytes = (b->l2_bytes ? b->l2_bytes : (size_t)128 << 20) * 2;
    VCL_CUDA(b, cudaMalloc(&b->flush_buf, b->flush_bytes));
  }
  VCL_CUDA(b, cudaMemsetAsync(b->flush_buf, 0, b->flush_bytes, b->stream));
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLBackendSetOption(ViennaCLBackend b, const char *name, long long value)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, name != nullptr, "null option name");
  if (!strcmp(name, "persistent_rows")) { b->persistent_rows = value; return ViennaCLSuccess; }
  if (!strcmp(name, "l2_resident")) { b->l2_resident = (int)value; return ViennaCLSuccess; }
  if (!strcmp(name, "persistent_cg_form")) { b->persistent_cg_form = (int)value; return ViennaCLSuccess; }
  return vcl_fail(b, ViennaCLB200InvalidArgument, "unknown option", __FILE__, __LINE__);
}

ViennaCLStatus ViennaCLBackendLaunchCount(ViennaCLBackend b, long long *launches)
{
  VCL_CHECK_BACKEND(b);
  *launches = b->launches;
  return ViennaCLSuccess;
}

// ---------------------------------------------------------------- row-block plans ----------------------------------------------------------------
} // extern "C"

__global__ void plan_check_kernel(const u32 *rp, int rows, const u32 *blk, int nb, int *bad)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= nb; i += gridDim.x * blockDim.x)
  {
    const u32 r0 = blk[i];
    bool wrong = r0 > (u32)rows || (i == 0 && r0 != 0u) || (i == nb && r0 != (u32)rows);
    if (!wrong && i < nb)
    {
      const u32 r1 = blk[i + 1];
      wrong = r1 < r0 || r1 > (u32)rows;
      if (!wrong && r1 - r0 > 1u)
        wrong = r1 - r0 > (u32)VCL_B200_CSR_BLOCK_ROWS || rp[r1] - rp[r0] > (u32)VCL_B200_CSR_BLOCK_NNZ;
    }
    if (wrong) *bad = 1;
  }
}

bool vcl_plan_ok(ViennaCLBackend b, const unsigned int *row_ptr, int rows, long long nnz, const unsigned int *row_blocks, int num_blocks)
{
  if (!row_blocks || num_blocks <= 0) return false;
  std::map<const void*, ViennaCLBackend_impl::PlanRec>::const_iterator it = b->plans.find(row_blocks);
  if (it != b->plans.end() && it->second.row_ptr == row_ptr && it->second.rows == rows && it->second.num_blocks == num_blocks) return it->second.ok;
  int *flag = reinterpret_cast<int*>(b->dscal + 48);
  int bad = 1;
  if (cudaMemsetAsync(flag, 0, sizeof(int), b->stream) == cudaSuccess)
  {
    plan_check_kernel<<<std::max(1, std::min((num_blocks + 256) / 256, b->sm_count * 4)), 256, 0, b->stream>>>(row_ptr, rows, row_blocks, num_blocks, flag);
    b->launches++;
    if (cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, b->stream) != cudaSuccess || cudaStreamSynchronize(b->stream) != cudaSuccess)
    { (void)cudaGetLastError(); bad = 1; }
  }
  if (nnz / num_blocks < 1200) bad = 1;                   // small blocks: the plan-free kernel is the faster one (common.cuh)
  const ViennaCLBackend_impl::PlanRec rec = {row_ptr, rows, num_blocks, bad == 0};
  b->plans[row_blocks] = rec;
  return rec.ok;
}

void vcl_plan_register(ViennaCLBackend b, const unsigned int *row_ptr, int rows, const unsigned int *row_blocks, int num_blocks)
{
  const ViennaCLBackend_impl::PlanRec rec = {row_ptr, rows, num_blocks, true};
  b->plans[row_blocks] = rec;
}

void vcl_plan_forget(ViennaCLBackend b, const void *dst, size_t bytes)
{
  if (b->plans.empty() || !dst) return;
  std::map<const void*, ViennaCLBackend_impl::PlanRec>::iterator lo = b->plans.lower_bound(dst);
  while (lo != b->plans.end() && (const char*)lo->first < (const char*)dst + (bytes ? bytes : 1)) lo = b->plans.erase(lo);
}

extern "C" {
// ---------------------------------------------------------------- memory ----------------------------------------------------------------
ViennaCLStatus ViennaCLCUDAMemAlloc(ViennaCLBackend b, void **ptr, size_t bytes)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, ptr != nullptr, "null out pointer");
  VCL_CUDA(b, cudaSetDevice(b->device));
  *ptr = nullptr;
  if (bytes == 0) return ViennaCLSuccess;
  VCL_CUDA(b, cudaMalloc(ptr, bytes));
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDAMemFree(ViennaCLBackend b, void *ptr)
{
  VCL_CHECK_BACKEND(b);
  if (!ptr) return ViennaCLSuccess;
  vcl_plan_forget(b, ptr, 0);
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  VCL_CUDA(b, cudaFree(ptr));
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDAMemWrite(ViennaCLBackend b, void *dst, size_t off, const void *src, size_t bytes, ViennaCLInt async)
{
  VCL_CHECK_BACKEND(b);
  if (bytes == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, dst && src, "null pointer");
  vcl_plan_forget(b, (char*)dst + off, bytes);
  VCL_CUDA(b, cudaMemcpyAsync((char*)dst + off, src, bytes, cudaMemcpyHostToDevice, b->stream));
  if (!async) VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDAMemRead(ViennaCLBackend b, const void *src, size_t off, void *dst, size_t bytes, ViennaCLInt async)
{
  VCL_CHECK_BACKEND(b);
  if (bytes == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, dst && src, "null pointer");
  VCL_CUDA(b, cudaMemcpyAsync(dst, (const char*)src + off, bytes, cudaMemcpyDeviceToHost, b->stream));
  if (!async) VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDAMemCopy(ViennaCLBackend b, const void *src, size_t soff, void *dst, size_t doff, size_t bytes)
{
  VCL_CHECK_BACKEND(b);
  if (bytes == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, dst && src, "null pointer");
  vcl_plan_forget(b, (char*)dst + doff, bytes);
  VCL_CUDA(b, cudaMemcpyAsync((char*)dst + doff, (const char*)src + soff, bytes, cudaMemcpyDeviceToDevice, b->stream));
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDAMemSet(ViennaCLBackend b, void *dst, ViennaCLInt value, size_t bytes)
{
  VCL_CHECK_BACKEND(b);
  if (bytes == 0) return ViennaCLSuccess;
  vcl_plan_forget(b, dst, bytes);
  VCL_CUDA(b, cudaMemsetAsync(dst, value, bytes, b->stream));
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLHostAllocPinned(ViennaCLBackend b, void **ptr, size_t bytes)
{
  VCL_CHECK_BACKEND(b);
  VCL_CUDA(b, cudaMallocHost(ptr, bytes ? bytes : 1));
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLHostFreePinned(ViennaCLBackend b, void *ptr)
{
  VCL_CHECK_BACKEND(b);
  if (ptr) VCL_CUDA(b, cudaFreeHost(ptr));
  return ViennaCLSuccess;
}

} // extern "C"
