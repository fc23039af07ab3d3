// launch.cuh -- grid sizing + launch of the SpMV kernel templates (shared by spmv.cu and solvers.cu).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include "spmv_kernels.cuh"

namespace VCL_NS
{

// Resident CTAs per SM for a kernel, queried once per KERNEL (all B200s of a box are identical).  The cache is keyed by the
// function address, not by its type: kernels that differ only in a non-type template argument (csr_stream_kernel<Epi, SPLIT>,
// sell_kernel<Epi, PERM>) share one function-pointer type, and each of them needs its own opt-in to > 48 KB of dynamic
// shared memory.
template<class K>
static int vcl_occupancy(K kernel, int threads, int dyn_smem = 0)
{
  static std::mutex mtx;
  static std::unordered_map<const void*, int> cache;
  std::lock_guard<std::mutex> lock(mtx);
  const void *key = reinterpret_cast<const void*>(kernel);
  std::unordered_map<const void*, int>::const_iterator it = cache.find(key);
  if (it != cache.end()) return it->second;
  int n = 0;
  if (dyn_smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, (size_t)dyn_smem) != cudaSuccess || n <= 0) n = 1;
  cache[key] = n;
  return n;
}

static inline bool vcl_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Grids are persistent: (resident CTAs per SM) x (SM count) CTAs loop over the row blocks, so the number of per-CTA
// reduction partials stays small and block b's partial is a deterministic function of the matrix alone.
template<class Epi>
static ViennaCLStatus vcl_launch_csr(ViennaCLBackend b, const ViennaCLCUDADcsr &A, XVec xv, Epi epi)
{
  CsrDev d = {A.rows, (u32)A.nnz, A.row_ptr, A.col_idx, A.values, A.row_blocks, A.row_blocks ? A.row_blocks + 1 : nullptr, A.num_blocks};
  if (A.row_blocks && A.num_blocks > 0 && vcl_aligned16(A.values) && vcl_aligned16(A.col_idx))
  {
    const int occ = vcl_occupancy(csr_stream_kernel<Epi, false>, CSR_BLOCK_THREADS, CSR_SMEM_BYTES);
    int grid = std::min(A.num_blocks, std::min(b->sm_count * occ, VCL_MAX_BLOCKS));
    csr_stream_kernel<Epi, false><<<grid, CSR_BLOCK_THREADS, CSR_SMEM_BYTES, b->stream>>>(d, xv, epi);
    VCL_LAUNCHED(b, "csr_stream_kernel");
  }
  else
  {
    const int occ = vcl_occupancy(csr_scalar_kernel<Epi>, 256);
    int grid = std::max(1, std::min(vcl_div_up(A.rows, 256), std::min(b->sm_count * occ, VCL_MAX_BLOCKS)));
    csr_scalar_kernel<Epi><<<grid, 256, 0, b->stream>>>(d, xv, epi);
    VCL_LAUNCHED(b, "csr_scalar_kernel");
  }
  return ViennaCLSuccess;
}

// Row-partitioned variant: a subset of the row blocks (interior or boundary list), x addressed as [owned | halo].
template<class Epi>
static ViennaCLStatus vcl_launch_csr_split(ViennaCLBackend b, const CsrDev &d, XVec xv, Epi epi, cudaStream_t stream)
{
  if (d.nblk <= 0) return ViennaCLSuccess;
  // Persistent by default.  VCL_B200_SPLIT_CHUNK=k (experiment knob) makes CTAs retire after ~k row blocks so that
  // communication kernels on the high-priority stream find SM slots while the interior blocks run.
  const int occ = vcl_occupancy(csr_stream_kernel<Epi, true>, CSR_BLOCK_THREADS, CSR_SMEM_BYTES);
  const int resident = b->sm_count * occ;
  static int chunk = -1;
  if (chunk < 0) { const char *e = getenv("VCL_B200_SPLIT_CHUNK"); chunk = e ? atoi(e) : 0; }
  int grid = std::min(d.nblk, std::min(resident, VCL_MAX_BLOCKS));
  if (chunk > 0) grid = std::min(d.nblk, std::min(std::max(resident, vcl_div_up(d.nblk, chunk)), VCL_MAX_BLOCKS));
  csr_stream_kernel<Epi, true><<<grid, CSR_BLOCK_THREADS, CSR_SMEM_BYTES, stream>>>(d, xv, epi);
  VCL_LAUNCHED(b, "csr_stream_kernel(split)");
  return ViennaCLSuccess;
}

template<class Epi>
static ViennaCLStatus vcl_launch_sell(ViennaCLBackend b, const ViennaCLCUDADsell &A, XVec xv, Epi epi)
{
  SellDev d = {A.rows, A.rows_per_block, A.columns_per_block, A.col_idx, A.block_start, A.values, A.row_perm};
  const int occ = A.row_perm ? vcl_occupancy(sell_kernel<Epi, true>, CSR_BLOCK_THREADS, CSR_SMEM_BYTES)
                             : vcl_occupancy(sell_kernel<Epi, false>, CSR_BLOCK_THREADS, CSR_SMEM_BYTES);
  const int C = A.rows_per_block;
  const int nslices = (A.rows - 1) / C + 1;
  const int spb = C <= CSR_BLOCK_THREADS ? CSR_BLOCK_THREADS / C : 1;
  int grid = std::max(1, std::min(vcl_div_up(nslices, spb), std::min(b->sm_count * occ, VCL_MAX_BLOCKS)));
  // SELL-C-sigma puts the widest slices at the start of every sorting window, i.e. at a fixed period of the block index: an odd
  // grid is coprime with that power-of-two period, so the wide blocks are dealt evenly to the persistent CTAs
  if (A.row_perm != nullptr && grid > 1 && (grid & 1) == 0) grid -= 1;
  if (A.row_perm) sell_kernel<Epi, true><<<grid, CSR_BLOCK_THREADS, CSR_SMEM_BYTES, b->stream>>>(d, xv, epi);
  else            sell_kernel<Epi, false><<<grid, CSR_BLOCK_THREADS, CSR_SMEM_BYTES, b->stream>>>(d, xv, epi);
  VCL_LAUNCHED(b, "sell_kernel");
  return ViennaCLSuccess;
}

template<class Epi>
static ViennaCLStatus vcl_launch_ell(ViennaCLBackend b, const ViennaCLCUDADhyb &A, XVec xv, Epi epi)
{
  EllDev d = {A.ell.rows, A.ell.internal_rows, A.ell.maxnnz, A.ell.coords, A.ell.elements, A.csr_rows, A.csr_cols, A.csr_elements};
  const int occ = vcl_occupancy(ell_kernel<Epi>, CSR_BLOCK_THREADS);
  int grid = std::max(1, std::min(vcl_div_up(A.ell.rows, CSR_BLOCK_THREADS), std::min(b->sm_count * occ, VCL_MAX_BLOCKS)));
  ell_kernel<Epi><<<grid, CSR_BLOCK_THREADS, 0, b->stream>>>(d, xv, epi);
  VCL_LAUNCHED(b, "ell_kernel");
  return ViennaCLSuccess;
}
} // namespace VCL_NS
