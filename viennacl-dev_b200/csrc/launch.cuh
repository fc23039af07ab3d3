// launch.cuh -- grid sizing + launch of the SpMV kernel templates (shared by spmv.cu and solvers.cu).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <map>
#include <utility>
#include "spmv_kernels.cuh"

namespace VCL_NS
{

// Resident CTAs per SM for a kernel on the handle's device, queried once per (DEVICE, KERNEL).  The opt-in to > 48 KB of
// dynamic shared memory (cudaFuncAttributeMaxDynamicSharedMemorySize) is a per-device attribute of the function, so a second
// handle on another device of the same process needs its own call.  The key uses the function ADDRESS, not its type:
// kernels that differ only in a non-type template argument (csr_stream_kernel<Epi, SPLIT>, sell_kernel<Epi, PERM>) share one
// function-pointer type.  The caller has made b->device current (VCL_CHECK_BACKEND).
template<class K>
static int vcl_occupancy(ViennaCLBackend b, K kernel, int threads, int dyn_smem = 0)
{
  static std::mutex mtx;
  static std::map<std::pair<int, const void*>, int> cache;
  std::lock_guard<std::mutex> lock(mtx);
  const std::pair<int, const void*> key(b->device, reinterpret_cast<const void*>(kernel));
  std::map<std::pair<int, const void*>, int>::const_iterator it = cache.find(key);
  if (it != cache.end()) return it->second;
  int n = 0;
  if (dyn_smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, (size_t)dyn_smem) != cudaSuccess || n <= 0) n = 1;
  cache[key] = n;
  return n;
}

// L2 policy for the streams of a CSR matrix (CsrDev::l2_mode).  Default: evict-first -- the matrix is streamed and must not push
// the gathered x entries out of L2.  Keeping a matrix that fits L2 resident (evict-last / normal) was MEASURED SLOWER on B200
// for BASELINE config 1 (1024^2 CG: 25.9 us per iteration evict-first, 27.8 evict-last, 28.5 normal; profiles/cg1024_l2_r2a.log):
// with evict-first the matrix streams from HBM while the vectors are served by L2, two sources in parallel.  Option
// "l2_resident" of the handle selects 1 (evict-last) or 2 (normal) for experiments.
static inline int vcl_l2_mode(ViennaCLBackend b, long long, long long)
{
  return b->l2_resident > 0 ? b->l2_resident : 0;
}

static inline bool vcl_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Grids are persistent: (resident CTAs per SM) x (SM count) CTAs loop over the row blocks, so the number of per-CTA
// reduction partials stays small and block b's partial is a deterministic function of the matrix alone.
template<class Epi>
static ViennaCLStatus vcl_launch_csr(ViennaCLBackend b, const ViennaCLCUDADcsr &A, XVec xv, Epi epi)
{
  CsrDev d = {A.rows, (u32)A.nnz, A.row_ptr, A.col_idx, A.values, A.row_blocks, A.row_blocks ? A.row_blocks + 1 : nullptr, A.num_blocks};
  d.l2_mode = vcl_l2_mode(b, A.nnz, A.rows);
  if (A.row_blocks && A.num_blocks > 0 && vcl_aligned16(A.values) && vcl_aligned16(A.col_idx) &&
      vcl_plan_ok(b, A.row_ptr, A.rows, A.nnz, A.row_blocks, A.num_blocks))
  {
    const int occ = vcl_occupancy(b, csr_stream_kernel<Epi, false>, CSR_BLOCK_THREADS, CSR_SMEM_BYTES);
    int grid = std::min(A.num_blocks, std::min(b->sm_count * occ, VCL_MAX_BLOCKS));
    csr_stream_kernel<Epi, false><<<grid, CSR_BLOCK_THREADS, CSR_SMEM_BYTES, b->stream>>>(d, xv, epi);
    VCL_LAUNCHED(b, "csr_stream_kernel");
  }
  else
  {
    const int occ = vcl_occupancy(b, csr_scalar_kernel<Epi>, 256);
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
  const int occ = vcl_occupancy(b, csr_stream_kernel<Epi, true>, CSR_BLOCK_THREADS, CSR_SMEM_BYTES);
  const int resident = b->sm_count * occ;
  static int chunk = -1;
  if (chunk < 0) { const char *e = getenv("VCL_B200_SPLIT_CHUNK"); chunk = e ? atoi(e) : 0; }
  int grid = std::min(d.nblk, std::min(resident, VCL_MAX_BLOCKS));
  if (chunk > 0) grid = std::min(d.nblk, std::min(std::max(resident, vcl_div_up(d.nblk, chunk)), VCL_MAX_BLOCKS));
  csr_stream_kernel<Epi, true><<<grid, CSR_BLOCK_THREADS, CSR_SMEM_BYTES, stream>>>(d, xv, epi);
  VCL_LAUNCHED(b, "csr_stream_kernel(split)");
  return ViennaCLSuccess;
}

template<class Epi, bool PERM, int CT>
static ViennaCLStatus vcl_launch_sell_as(ViennaCLBackend b, const SellDev &d, XVec xv, Epi epi)
{
  const int occ = vcl_occupancy(b, sell_kernel<Epi, PERM, CT>, CSR_BLOCK_THREADS, CSR_SMEM_BYTES);
  const int C = d.C;
  const int nslices = (d.rows - 1) / C + 1;
  const int spb = C <= CSR_BLOCK_THREADS ? CSR_BLOCK_THREADS / C : 1;
  int grid = std::max(1, std::min(vcl_div_up(nslices, spb), std::min(b->sm_count * occ, VCL_MAX_BLOCKS)));
  // SELL-C-sigma puts the widest slices at the start of every sorting window, i.e. at a fixed period of the block index: an odd
  // grid is coprime with that power-of-two period, so the wide blocks are dealt evenly to the persistent CTAs
  if (PERM && grid > 1 && (grid & 1) == 0) grid -= 1;
  sell_kernel<Epi, PERM, CT><<<grid, CSR_BLOCK_THREADS, CSR_SMEM_BYTES, b->stream>>>(d, xv, epi);
  VCL_LAUNCHED(b, "sell_kernel");
  return ViennaCLSuccess;
}

// Row-partitioned slab in SELL (dist.cu): x = [owned | halo], halo control inside d; sigma = 1 only.
template<class Epi>
static ViennaCLStatus vcl_launch_sell_split(ViennaCLBackend b, const SellDev &d, XVec xv, Epi epi)
{
  const int C = d.C;
  const int nslices = (d.rows - 1) / C + 1;
  const int spb = C <= CSR_BLOCK_THREADS ? CSR_BLOCK_THREADS / C : 1;
  if (C == 32)
  {
    const int occ = vcl_occupancy(b, sell_kernel<Epi, false, 32, true>, CSR_BLOCK_THREADS, CSR_SMEM_BYTES);
    const int grid = std::max(1, std::min(vcl_div_up(nslices, spb), std::min(b->sm_count * occ, VCL_MAX_BLOCKS)));
    sell_kernel<Epi, false, 32, true><<<grid, CSR_BLOCK_THREADS, CSR_SMEM_BYTES, b->stream>>>(d, xv, epi);
  }
  else
  {
    const int occ = vcl_occupancy(b, sell_kernel<Epi, false, 0, true>, CSR_BLOCK_THREADS, CSR_SMEM_BYTES);
    const int grid = std::max(1, std::min(vcl_div_up(nslices, spb), std::min(b->sm_count * occ, VCL_MAX_BLOCKS)));
    sell_kernel<Epi, false, 0, true><<<grid, CSR_BLOCK_THREADS, CSR_SMEM_BYTES, b->stream>>>(d, xv, epi);
  }
  VCL_LAUNCHED(b, "sell_kernel(split)");
  return ViennaCLSuccess;
}

template<class Epi>
static ViennaCLStatus vcl_launch_sell(ViennaCLBackend b, const ViennaCLCUDADsell &A, XVec xv, Epi epi)
{
  SellDev d = {A.rows, A.rows_per_block, A.columns_per_block, A.col_idx, A.block_start, A.values, A.row_perm,
               nullptr, 0u, nullptr, 0ULL, nullptr, nullptr, nullptr, nullptr};
  // the slice height of the reference's default layout gets its own instantiation (sliced_ell_matrix.hpp:146-147: C = 32)
  if (A.row_perm) return A.rows_per_block == 32 ? vcl_launch_sell_as<Epi, true, 32>(b, d, xv, epi) : vcl_launch_sell_as<Epi, true, 0>(b, d, xv, epi);
  return A.rows_per_block == 32 ? vcl_launch_sell_as<Epi, false, 32>(b, d, xv, epi) : vcl_launch_sell_as<Epi, false, 0>(b, d, xv, epi);
}

template<class Epi>
static ViennaCLStatus vcl_launch_ell(ViennaCLBackend b, const ViennaCLCUDADhyb &A, XVec xv, Epi epi)
{
  EllDev d = {A.ell.rows, A.ell.internal_rows, A.ell.maxnnz, A.ell.coords, A.ell.elements, A.csr_rows, A.csr_cols, A.csr_elements};
  const int occ = vcl_occupancy(b, ell_kernel<Epi>, CSR_BLOCK_THREADS);
  int grid = std::max(1, std::min(vcl_div_up(A.ell.rows, CSR_BLOCK_THREADS), std::min(b->sm_count * occ, VCL_MAX_BLOCKS)));
  ell_kernel<Epi><<<grid, CSR_BLOCK_THREADS, 0, b->stream>>>(d, xv, epi);
  VCL_LAUNCHED(b, "ell_kernel");
  return ViennaCLSuccess;
}
} // namespace VCL_NS
