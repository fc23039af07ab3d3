// spmv_kernels.cuh -- sm_100a SpMV kernels for CSR (row-block streaming) and SELL-C-sigma, parameterised by an
// epilogue functor so that the same streaming code serves  y = alpha*A*x + beta*y  (prod_impl) and the fused
// "SpMV + inner products" steps of the pipelined solvers.
//
// Replaces cuda/sparse_matrix_operations.hpp:137-249 (K1/K2), :2196-2237 (K4) and the SpMV halves of
// cuda/iterative_operations.hpp:113-274, :543-607, :895-1074, :1381-1453 (K6-K8, K11-K13).
//
// Data path (CSR): each CTA owns whole rows [blk[b], blk[b+1]) (<= 256 rows, <= 2048 nnz).  The contiguous value and
// column-index ranges of those rows are streamed global->shared with 16-byte cp.async (no register staging, L1 bypass),
// then one thread per row walks its entries in shared memory IN STORAGE ORDER with a single accumulation chain -- the same
// operation order and roundings as the reference host backend (host_based/sparse_matrix_operations.hpp:167-184), so
// results agree bit-for-bit with it (see madd()).  x is gathered through L1/L2 (adjacent rows of a stencil matrix
// gather adjacent x entries, so the gathers of a warp coalesce).
#pragma once
#include "common.cuh"
#include "solver_state.cuh"
#include "peer.cuh"

#define CSR_BLOCK_THREADS 256
#define CSR_CAP (VCL_B200_CSR_BLOCK_NNZ)          // staged non-zeros per row block
#define CSR_STAGE (CSR_CAP + 4)                   // + alignment slack

struct CsrDev
{
  int rows; u32 nnz;
  const u32 *rp, *ci; const double *va;
  const u32 *blk; int nblk;
  const u32 *blk_list;     // optional indirection: process row blocks blk_list[0..nblk) (interior / boundary subsets)
  // peer-memory halo (row-partitioned path, peer.cuh): list positions >= wait_from read halo columns and must first see
  // wait_flags[q] >= wait_seq for every source rank q in wait_mask.  wait_mask == 0: nothing to wait for.
  int wait_from; unsigned int wait_mask; const unsigned long long *wait_flags; unsigned long long wait_seq; int *err;
  unsigned long long *dbg; unsigned long long dbg_seq;     // VCL_PEER_DEBUG builds only
};

struct SellDev
{
  int rows; int C;
  const u32 *cpb, *ci, *bs; const double *va;
};

// x operand.  Row-partitioned matrices address [owned | halo]: columns >= split are read from x2 (the halo receive buffer).
struct XVec { const double *x; int off, inc; const double *x2; u32 split; };

// SPLIT: one load from a selected base (no branch).  Halo entries are written by peer GPUs; reading them through L1 is safe
// because a CTA touches the halo only after its acquire on the arrival flag (peer.cuh) and L1 does not outlive a launch.
template<bool SPLIT>
__device__ __forceinline__ double xload(const XVec &xv, u32 c)
{
  if (SPLIT)
  {
    const double *base = (c >= xv.split) ? (xv.x2 - xv.split) : xv.x;
    return base[c];
  }
  return xv.x[(size_t)c * xv.inc + xv.off];
}

// In-row CSR accumulation step.  The reference host backend (host_based/sparse_matrix_operations.hpp:167-184), built with
// g++ -O3 for x86-64-v3, evaluates `dot += a*x` as a rounded multiply followed by a rounded add (GCC does not form FMA
// chains in reductions under generic tuning); the same two roundings are used here so that CSR results match it bit for bit.
__device__ __forceinline__ double madd(double a, double x, double acc) { return __dadd_rn(acc, __dmul_rn(a, x)); }

// The matrix arrays are read exactly once per product: they are streamed with an L2 evict-first policy so that they do not
// push the gathered x entries (re-used by the rows of the next planes, tens of MB of streamed matrix data later) out of L2.
__device__ __forceinline__ unsigned long long l2_evict_first_policy()
{
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, unsigned long long pol)
{
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" :: "r"(s), "l"(gmem_src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template<int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------
// Epilogues
// ------------------------------------------------------------------------------------------------
// y[off + r*inc] = alpha*dot + beta*y   (spmv_alpha_beta semantics, cuda/sparse_matrix_operations.hpp:130: beta == 0 -> y not read)
struct EpiAxpby
{
  double *y; int off, inc; double alpha, beta;
  static constexpr int NQ = 0;
  __device__ __forceinline__ bool skip() const { return false; }
  // pre(): the epilogue's own per-row operand, requested BEFORE the row's gather chain so that its latency overlaps
  __device__ __forceinline__ double pre(u32 r) const
  {
    return (beta != 0.0) ? y[(size_t)r * (size_t)inc + (size_t)off] : 0.0;
  }
  __device__ __forceinline__ void row(u32 r, double dot, double y_old)
  {
    size_t idx = (size_t)r * (size_t)inc + (size_t)off;
    // same operations as the reference host build: t = alpha*dot (rounded), then one fused beta*y + t
    if (beta != 0.0) y[idx] = fma(beta, y_old, __dmul_rn(alpha, dot));
    else             y[idx] = __dmul_rn(alpha, dot);
  }
  __device__ __forceinline__ void finish(double *) {}
};

// ------------------------------------------------------------------------------------------------
// CSR, row-block streaming
// ------------------------------------------------------------------------------------------------
template<class Epi, bool SPLIT>
__global__ void __launch_bounds__(CSR_BLOCK_THREADS, 8)
csr_stream_kernel(CsrDev A, XVec xv, Epi epi)
{
  __shared__ __align__(16) double s_val[CSR_STAGE];
  __shared__ __align__(16) u32    s_col[CSR_STAGE];
  __shared__ u32 s_rp[CSR_BLOCK_THREADS + 1];
  __shared__ double s_red[(Epi::NQ > 0 ? Epi::NQ : 1) * 32];

  if (epi.skip()) return;
  const int tid = threadIdx.x;
  const unsigned long long pol = l2_evict_first_policy();
#ifdef VCL_PEER_DEBUG
  if (SPLIT && A.dbg && blockIdx.x == 0 && tid == 0) A.dbg[(A.dbg_seq % 1024) * 4 + 2] = global_ns();
#endif
  bool halo_ready = !(SPLIT && A.wait_mask != 0u);

  for (int bi = blockIdx.x; bi < A.nblk; bi += gridDim.x)
  {
    if (SPLIT && !halo_ready && bi >= A.wait_from)
    {
      // boundary row blocks come last in the list: by now the neighbours' pushes have normally landed
#ifdef VCL_PEER_DEBUG
      const u64 t_w = global_ns();
#endif
      if (tid < VCL_MAX_PEERS && ((A.wait_mask >> tid) & 1u)) peer_wait(A.wait_flags + tid, A.wait_seq, A.err);
      __syncthreads();
#ifdef VCL_PEER_DEBUG
      if (A.dbg && tid == 0) atomicMax(A.dbg + (A.dbg_seq % 1024) * 4 + 3, global_ns() - t_w);
#endif
      halo_ready = true;
    }
    const u32 b = A.blk_list ? A.blk_list[bi] : (u32)bi;
    const u32 r0 = A.blk[b], r1 = A.blk[b + 1];
    const u32 nrows = r1 - r0;
    const u32 n0 = A.rp[r0], n1 = A.rp[r1];

    if (n1 - n0 > CSR_CAP)
    {
      // one long row: the whole CTA strides over it (summation order differs from the sequential reference; tolerance-level parity)
      double part[1] = {0.0};
      for (u32 k = n0 + tid; k < n1; k += CSR_BLOCK_THREADS)
        part[0] = fma(A.va[k], xload<SPLIT>(xv, A.ci[k]), part[0]);
      __shared__ double s_long[32];
      block_sum<1>(part, s_long);
      if (tid == 0) epi.row(r0, part[0], epi.pre(r0));
      __syncthreads();
      continue;
    }

    // ---- stage values / column indices: 16-byte cp.async from the enclosing 16-byte aligned range ----
    const u32 a0 = n0 & ~3u;
    const u32 cnt = n1 - a0;
    for (u32 i = tid * 2; i < cnt; i += CSR_BLOCK_THREADS * 2)
    {
      if (a0 + i + 2 <= A.nnz) cp_async16(&s_val[i], A.va + a0 + i, pol);
      else if (a0 + i < A.nnz) s_val[i] = A.va[a0 + i];
    }
    for (u32 i = tid * 4; i < cnt; i += CSR_BLOCK_THREADS * 4)
    {
      if (a0 + i + 4 <= A.nnz) cp_async16(&s_col[i], A.ci + a0 + i, pol);
      else { for (u32 k = 0; k < 4 && a0 + i + k < A.nnz; ++k) s_col[i + k] = A.ci[a0 + i + k]; }
    }
    cp_async_commit();
    if ((u32)tid <= nrows) s_rp[tid] = A.rp[r0 + tid];
    if (tid == 0 && nrows == CSR_BLOCK_THREADS) s_rp[CSR_BLOCK_THREADS] = n1;
    const double pre = ((u32)tid < nrows) ? epi.pre(r0 + tid) : 0.0;      // in flight together with the staging copies
    cp_async_wait<0>();
    __syncthreads();

    // ---- one thread per row, sequential fma chain in storage order ----
    if ((u32)tid < nrows)
    {
      u32 j = s_rp[tid] - a0;
      const u32 e = s_rp[tid + 1] - a0;
      double dot = 0.0;
      for (; j + 4 <= e; j += 4)
      {
        const u32 c0 = s_col[j], c1 = s_col[j + 1], c2 = s_col[j + 2], c3 = s_col[j + 3];
        const double x0 = xload<SPLIT>(xv, c0), x1 = xload<SPLIT>(xv, c1);
        const double x2 = xload<SPLIT>(xv, c2), x3 = xload<SPLIT>(xv, c3);
        dot = madd(s_val[j], x0, dot); dot = madd(s_val[j + 1], x1, dot);
        dot = madd(s_val[j + 2], x2, dot); dot = madd(s_val[j + 3], x3, dot);
      }
      if (j + 2 <= e)
      {
        const u32 c0 = s_col[j], c1 = s_col[j + 1];
        const double x0 = xload<SPLIT>(xv, c0), x1 = xload<SPLIT>(xv, c1);
        dot = madd(s_val[j], x0, dot); dot = madd(s_val[j + 1], x1, dot);
        j += 2;
      }
      if (j < e) dot = madd(s_val[j], xload<SPLIT>(xv, s_col[j]), dot);
      epi.row(r0 + tid, dot, pre);
    }
    __syncthreads();
  }
  epi.finish(s_red);
}

// Plan-free CSR kernel: one thread per row, sequential order (used when no row blocks are supplied or the
// arrays are not 16-byte aligned, e.g. oddly offset user-wrapped buffers).
template<class Epi>
__global__ void __launch_bounds__(256)
csr_scalar_kernel(CsrDev A, XVec xv, Epi epi)
{
  __shared__ double s_red[(Epi::NQ > 0 ? Epi::NQ : 1) * 32];
  if (epi.skip()) return;
  const double * __restrict__ x = xv.x;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < A.rows; r += (long long)gridDim.x * blockDim.x)
  {
    double dot = 0.0;
    const u32 e = A.rp[r + 1];
    for (u32 k = A.rp[r]; k < e; ++k)
      dot = madd(A.va[k], x[(size_t)A.ci[k] * xv.inc + xv.off], dot);
    epi.row((u32)r, dot, epi.pre((u32)r));
  }
  epi.finish(s_red);
}

// ------------------------------------------------------------------------------------------------
// SELL-C-sigma (sigma = 1).  Slices are stored back to back, so the slices of a CTA (256/C of them: 8 for the default
// C = 32) form ONE contiguous range of values and of column indices: it is streamed global->shared with 16-byte cp.async
// exactly like a CSR row block, then one thread per row walks its slice-column-major entries (stride C in shared memory:
// consecutive rows hit consecutive banks) with one fma chain.  Zero-valued (padding) slots never touch x
// (cuda/sparse_matrix_operations.hpp:2231, host :1833).  Slices wider than the staging buffer take the direct path.
// The multiply-adds are fused: that is what the reference host build does for SELL (oracle/vcl_oracle.c, ARITHMETIC).
// ------------------------------------------------------------------------------------------------
template<class Epi>
__global__ void __launch_bounds__(CSR_BLOCK_THREADS, 8)
sell_kernel(SellDev A, XVec xv, Epi epi)
{
  __shared__ __align__(16) double s_val[CSR_STAGE];
  __shared__ __align__(16) u32    s_col[CSR_STAGE];
  __shared__ double s_red[(Epi::NQ > 0 ? Epi::NQ : 1) * 32];
  if (epi.skip()) return;
  const double * __restrict__ x = xv.x;
  const double * __restrict__ va = A.va;
  const u32 * __restrict__ ci = A.ci;
  const int tid = threadIdx.x;
  const unsigned long long pol = l2_evict_first_policy();
  const u32 C = (u32)A.C;
  const u32 nslices = (u32)((A.rows - 1) / A.C + 1);
  const u32 spb = C <= CSR_BLOCK_THREADS ? CSR_BLOCK_THREADS / C : 1u;      // slices per CTA pass
  const u32 nblocks = (nslices + spb - 1) / spb;
  const bool can_stage = (C % 4u) == 0u && C <= CSR_BLOCK_THREADS &&
                         ((reinterpret_cast<uintptr_t>(va) | reinterpret_cast<uintptr_t>(ci)) & 15u) == 0u;

  for (u32 b = blockIdx.x; b < nblocks; b += gridDim.x)
  {
    const u32 s0 = b * spb, s1 = min(s0 + spb, nslices);
    const u32 base = A.bs[s0];
    const u32 end = A.bs[s1 - 1] + A.cpb[s1 - 1] * C;
    const u32 cnt = end - base;
    const bool staged = can_stage && cnt <= CSR_CAP;
    if (staged)
    {
      // base is a multiple of C (hence of 4) and cnt a multiple of C: whole 16-byte packets, never past the arrays
      for (u32 i = tid * 2; i < cnt; i += CSR_BLOCK_THREADS * 2) cp_async16(&s_val[i], va + base + i, pol);
      for (u32 i = tid * 4; i < cnt; i += CSR_BLOCK_THREADS * 4) cp_async16(&s_col[i], ci + base + i, pol);
      cp_async_commit();
      // (s1 - s0) * C <= 256 here: one row per thread
      const bool active = (u32)tid < (s1 - s0) * C;
      u32 w = 0, idx = 0;
      long long r = 0;
      if (active)
      {
        const u32 slice = s0 + tid / C;
        r = (long long)slice * C + (tid % C);
        w = A.cpb[slice];
        idx = A.bs[slice] + (tid % C) - base;
      }
      const double pre = (active && r < A.rows) ? epi.pre((u32)r) : 0.0;
      cp_async_wait<0>();
      __syncthreads();
      if (active && r < A.rows)
      {
        double acc = 0.0;
        u32 j = 0;
        for (; j + 4 <= w; j += 4, idx += 4 * C)
        {
          const double v0 = s_val[idx], v1 = s_val[idx + C], v2 = s_val[idx + 2 * C], v3 = s_val[idx + 3 * C];
          const u32 c0 = s_col[idx], c1 = s_col[idx + C], c2 = s_col[idx + 2 * C], c3 = s_col[idx + 3 * C];
          const double x0 = (v0 != 0.0) ? x[(size_t)c0 * xv.inc + xv.off] : 0.0;
          const double x1 = (v1 != 0.0) ? x[(size_t)c1 * xv.inc + xv.off] : 0.0;
          const double x2 = (v2 != 0.0) ? x[(size_t)c2 * xv.inc + xv.off] : 0.0;
          const double x3 = (v3 != 0.0) ? x[(size_t)c3 * xv.inc + xv.off] : 0.0;
          if (v0 != 0.0) acc = fma(x0, v0, acc);
          if (v1 != 0.0) acc = fma(x1, v1, acc);
          if (v2 != 0.0) acc = fma(x2, v2, acc);
          if (v3 != 0.0) acc = fma(x3, v3, acc);
        }
        for (; j < w; ++j, idx += C)
        {
          const double v0 = s_val[idx];
          if (v0 != 0.0) acc = fma(x[(size_t)s_col[idx] * xv.inc + xv.off], v0, acc);
        }
        epi.row((u32)r, acc, pre);
      }
      __syncthreads();
    }
    else
    {
      // direct path (very wide slices, C not a multiple of 4, or C > 256): coalesced 8-/4-byte loads straight from global
      for (u32 t = tid; t < (s1 - s0) * C; t += CSR_BLOCK_THREADS)
      {
        const u32 slice = s0 + t / C;
        const long long r = (long long)slice * C + (t % C);
        if (r >= A.rows) continue;
        const u32 w = A.cpb[slice];
        size_t idx = (size_t)A.bs[slice] + (t % C);
        double acc = 0.0;
        u32 j = 0;
        for (; j + 4 <= w; j += 4, idx += 4 * (size_t)C)
        {
          const double v0 = va[idx], v1 = va[idx + C], v2 = va[idx + 2 * (size_t)C], v3 = va[idx + 3 * (size_t)C];
          const u32 c0 = ci[idx], c1 = ci[idx + C], c2 = ci[idx + 2 * (size_t)C], c3 = ci[idx + 3 * (size_t)C];
          const double x0 = (v0 != 0.0) ? x[(size_t)c0 * xv.inc + xv.off] : 0.0;
          const double x1 = (v1 != 0.0) ? x[(size_t)c1 * xv.inc + xv.off] : 0.0;
          const double x2 = (v2 != 0.0) ? x[(size_t)c2 * xv.inc + xv.off] : 0.0;
          const double x3 = (v3 != 0.0) ? x[(size_t)c3 * xv.inc + xv.off] : 0.0;
          if (v0 != 0.0) acc = fma(x0, v0, acc);
          if (v1 != 0.0) acc = fma(x1, v1, acc);
          if (v2 != 0.0) acc = fma(x2, v2, acc);
          if (v3 != 0.0) acc = fma(x3, v3, acc);
        }
        for (; j < w; ++j, idx += C)
        {
          const double v0 = va[idx];
          if (v0 != 0.0) acc = fma(x[(size_t)ci[idx] * xv.inc + xv.off], v0, acc);
        }
        epi.row((u32)r, acc, epi.pre((u32)r));
      }
    }
  }
  epi.finish(s_red);
}
