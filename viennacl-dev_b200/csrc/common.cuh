// common.cuh -- shared internals of libvcl_b200.so (backend handle, error plumbing, deterministic reductions).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <map>
#include <string>
#include <vector>
#include "vcl_b200.h"

typedef unsigned int u32;

// ------------------------------------------------------------------------------------------------
// Backend handle (the "stream and/or device descriptors" libviennacl left as a TODO,
// libviennacl/src/viennacl_private.hpp:38-41).  A handle is single-threaded; handles are independent.
// ------------------------------------------------------------------------------------------------

struct ViennaCLBackend_impl
{
  int device = 0;
  int sm_count = 148;
  size_t l2_bytes = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  cudaStream_t comm_stream = nullptr;          // halo traffic (multi-GPU)
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;  // generic fork/join events
  cudaEvent_t tm_begin = nullptr, tm_end = nullptr;
  std::string last_error;
  long long launches = 0;

  // reduction scratch shared by all kernels launched through this handle (kernels on one stream never overlap)
  double *partials = nullptr;                  // [VCL_MAX_QUANT][VCL_MAX_BLOCKS]
  unsigned int *tickets = nullptr;             // [16], zero between kernels
  double *dscal = nullptr;                     // VCL_DSCAL_COUNT device doubles: results of stand-alone reductions etc. (layout below)
  double *hscal = nullptr;                     // pinned mirror of dscal
  void *dstate = nullptr;                      // device solver state (SolverState of the build that runs, solver_state.cuh)
  void *hstate = nullptr;                      // pinned mirror
  void *flush_buf = nullptr; size_t flush_bytes = 0;

  // workspace pool for solver temporaries (grown on demand, freed at destroy)
  void *ws = nullptr; size_t ws_bytes = 0;

  // NCCL (dlopen'ed lazily, dist.cu)
  void *nccl_comm = nullptr;
  int rank = 0, world = 1;

  // per-handle options (ViennaCLBackendSetOption) and device facts
  long long persistent_rows = -1;              // row limit of the persistent cooperative solver kernels; -1: built-in default, 0: never
  int l2_resident = -1;                        // keep small matrices resident in L2 inside the persistent kernels: -1 auto, 0 never, 1 always
  int persistent_cg_form = 0;                  // 0: by size, 1 / 3: one-pass persistent CG (one grid barrier per iteration; 2 / 3 CTAs per SM), 2: the two-phase form
  int coop_launch = 0;                         // device supports cooperative launches

  // CSR row-block plans this handle has seen (key: device address of the plan), see vcl_plan_ok
  struct PlanRec { const void *row_ptr; int rows, num_blocks; bool ok; };
  std::map<const void*, PlanRec> plans;
};

// Is (row_blocks, num_blocks) a plan the TMA row-block kernels may use for this matrix -- every block <= VCL_B200_CSR_BLOCK_ROWS rows
// and <= VCL_B200_CSR_BLOCK_NNZ entries, or one longer row?  Plans made by ViennaCLCUDAcsr_row_blocks are registered as such; any
// other plan (e.g. the reference's own handle3() blocks, compressed_matrix.hpp:1152-1188: <= 1024 entries but any number of
// rows) is checked ONCE on the device (one small kernel + one 4-byte read) and the verdict is cached per plan address; a plan
// that breaks the limits sends the product to the plan-free kernel instead of giving wrong rows -- and so does a valid foreign
// plan whose blocks are small (fewer than 1200 entries on average: the staging buffers stay half empty; measured at 256^3 with
// the reference's 1024-entry blocks: 0.329 ms through the TMA kernel, 0.286 ms plan-free, profiles/ab_plans_r2.log).  Writing to /
// freeing the plan's memory through the C-ABI drops the cached verdict (vcl_plan_forget).
bool vcl_plan_ok(ViennaCLBackend b, const unsigned int *row_ptr, int rows, long long nnz, const unsigned int *row_blocks, int num_blocks);
void vcl_plan_register(ViennaCLBackend b, const unsigned int *row_ptr, int rows, const unsigned int *row_blocks, int num_blocks);
void vcl_plan_forget(ViennaCLBackend b, const void *dst, size_t bytes);

// Layout of dscal (slots of 8 bytes, whichever precision runs): [0, 16) solver set-up reductions, [16, 32) per-op API chunk sums,
// [32, 40) row-partitioned CG rank-local sums, [40, 44) coo2csr flag, [44, 48) row-block scratch, [48, 49) plan check, [64, 64 + VCL_GMRES_MAX_KRYLOV)
// folded <v_i, v_k> of the per-op Gram-Schmidt stage 2.
#define VCL_DSCAL_COUNT 192
#define VCL_DSCAL_GS_FOLD 64

#define VCL_MAX_BLOCKS 16384    // upper bound on the grid of any reducing kernel (partials: 64 x 16384 doubles = 8 MB)
#define VCL_MAX_QUANT  64       // quantities reduced at once (GMRES stage 1 reduces up to VCL_GMRES_MAX_KRYLOV dots)

ViennaCLStatus vcl_fail(ViennaCLBackend b, ViennaCLStatus st, const char *what, const char *file, int line);
ViennaCLStatus vcl_cuda_fail(ViennaCLBackend b, cudaError_t e, const char *what, const char *file, int line);
ViennaCLStatus vcl_ws_reserve(ViennaCLBackend b, size_t bytes);   // ensures b->ws holds >= bytes

// every entry point: a handle is bound to ONE device; make it current for the calling thread (several handles on several
// devices may live in one process, include/viennacl/backend/mem_handle.hpp)
#define VCL_CHECK_BACKEND(b) do { if (!(b)) return ViennaCLB200NotInitialized; (void)cudaSetDevice((b)->device); } while (0)
#define VCL_REQUIRE(b, cond, msg) do { if (!(cond)) return vcl_fail((b), ViennaCLB200InvalidArgument, msg, __FILE__, __LINE__); } while (0)
#define VCL_CUDA(b, expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return vcl_cuda_fail((b), e__, #expr, __FILE__, __LINE__); } while (0)
#define VCL_TRY(expr) do { ViennaCLStatus s__ = (expr); if (s__ != ViennaCLSuccess) return s__; } while (0)
// after every launch (cf. VIENNACL_CUDA_LAST_ERROR_CHECK, linalg/cuda/common.hpp:30,165-175)
#define VCL_LAUNCHED(b, name) do { (b)->launches++; cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return vcl_cuda_fail((b), e__, name, __FILE__, __LINE__); } while (0)

static inline int vcl_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// Device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

template<typename T>
__device__ __forceinline__ T warp_sum(T v)
{
  // XOR butterfly: every lane ends with the same, order-deterministic value.
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// Sum of NQ per-thread values over the block; result valid in thread 0.  smem: NQ * 32 doubles.
template<int NQ, typename T>
__device__ __forceinline__ void block_sum(T (&v)[NQ], T *smem)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int q = 0; q < NQ; ++q) v[q] = warp_sum(v[q]);
  __syncthreads();               // smem may still be in use by the caller
  if (lane == 0)
  {
#pragma unroll
    for (int q = 0; q < NQ; ++q) smem[q * 32 + wid] = v[q];
  }
  __syncthreads();
  if (wid == 0)
  {
#pragma unroll
    for (int q = 0; q < NQ; ++q)
    {
      T t = (lane < nw) ? smem[q * 32 + lane] : T(0);
      v[q] = warp_sum(t);
    }
  }
}

// Grid-wide deterministic reduction, second stage done by whichever block finishes last (fixed summation order, so the
// result does not depend on which block that is).  Every block calls this with its NQ block-local sums (valid in thread 0
// after block_sum).  Returns true in ALL threads of the last block, with totals[] valid in thread 0 of that block.
// partials: [NQ][VCL_MAX_BLOCKS]; ticket: one counter, left at zero on exit.
template<int NQ, typename T>
__device__ __forceinline__ bool grid_sum_last_block(T (&v)[NQ], T *partials, unsigned int *ticket, T *smem)
{
  __shared__ bool s_last;
  block_sum<NQ, T>(v, smem);
  if (threadIdx.x == 0)
  {
#pragma unroll
    for (int q = 0; q < NQ; ++q) partials[q * VCL_MAX_BLOCKS + blockIdx.x] = v[q];
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  T acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q)
  {
    acc[q] = T(0);
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x)
      acc[q] += __ldcg(partials + q * VCL_MAX_BLOCKS + i);
  }
  block_sum<NQ, T>(acc, smem);
  if (threadIdx.x == 0)
  {
#pragma unroll
    for (int q = 0; q < NQ; ++q) v[q] = acc[q];
    *ticket = 0u;
  }
  return true;
}

#endif // __CUDACC__
