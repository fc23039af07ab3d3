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
#include <type_traits>
#include "common.cuh"
#include "prec.cuh"
#include "solver_state.cuh"
#include "peer.cuh"

namespace VCL_NS
{

#ifndef CSR_BLOCK_THREADS
#define CSR_BLOCK_THREADS 256
#endif
#define CSR_CAP (VCL_B200_CSR_BLOCK_NNZ)          // staged non-zeros per row block
#define CSR_STAGE (CSR_CAP + 4)                   // + alignment slack
#define CSR_LONG_ROW 64                           // rows longer than this are summed by a whole warp (see csr_stream_body)

struct CsrDev
{
  int rows; u32 nnz;
  const u32 *rp, *ci; const real *va;
  const u32 *blk_start, *blk_end; int nblk;     // row block i covers rows [blk_start[i], blk_end[i]); for a plain plan
                                                // blk_end == blk_start + 1 (compressed_matrix handle3() layout)
  // peer-memory halo (row-partitioned path, peer.cuh): list positions >= wait_from read halo columns and must first see
  // wait_flags[q] >= wait_seq for every source rank q in wait_mask.  wait_mask == 0: nothing to wait for.
  int wait_from; unsigned int wait_mask; const unsigned long long *wait_flags; unsigned long long wait_seq; int *err;
  // halo push fused into the head of the product launch (plain row-partitioned products): every CTA first sends its share of
  // this rank's boundary entries of x to the neighbours (push != NULL; sequence number = wait_seq), then joins the row-block loop
  const HaloPush *push; const u32 *push_idx; unsigned int *push_ticket;
  int l2_mode;                                  // L2 policy of the matrix streams: 0 evict-first (streamed once), 1 evict-last (a matrix that
                                                // fits L2 and is re-read every solver iteration), 2 normal
  unsigned long long *dbg; unsigned long long dbg_seq;     // VCL_PEER_DEBUG builds only
};

struct SellDev
{
  int rows; int C;
  const u32 *cpb, *ci, *bs; const real *va;
  const u32 *perm;              // SELL-C-sigma: storage row -> matrix row (0xFFFFFFFF: padding); NULL: identity
  // row-partitioned slabs (sell_kernel<..., SPLIT = true>, dist.cu): x is addressed as [owned | halo]; a CTA waits for the neighbours'
  // halo flags before its first pass that gathers from the halo (needs[pass] != 0); push != NULL: the first CTAs send this rank's
  // boundary entries of x from the head of the kernel (halo_push_share), as in the CSR kernel
  const unsigned char *needs; unsigned int wait_mask; const unsigned long long *wait_flags; unsigned long long wait_seq; int *err;
  const HaloPush *push; const u32 *push_idx; unsigned int *push_ticket;
};

// x operand.  Row-partitioned matrices address [owned | halo]: columns >= split are read from x2 (the halo receive buffer).
// `x` already points at the first entry (start offset folded in); inc8 = stride in BYTES, so an entry address is ONE
// 32x32+64-bit multiply-add.  Build with make_xvec().
struct XVec { const real *x; u32 inc8; const real *x2; u32 split; };

static inline XVec make_xvec(const real *x, long long off, long long inc, const real *x2 = nullptr, u32 split = 0)
{
  XVec v = {x + off, (u32)(inc * (long long)sizeof(real)), x2, split};
  return v;
}

// SPLIT: one load from a selected base (no branch).  Halo entries are written by peer GPUs; reading them through L1 is safe
// because a CTA touches the halo only after its acquire on the arrival flag (peer.cuh) and L1 does not outlive a launch.
template<bool SPLIT>
__device__ __forceinline__ real xload(const XVec &xv, u32 c)
{
  if (SPLIT)
  {
    const real *base = (c >= xv.split) ? (xv.x2 - xv.split) : xv.x;
    return base[c];
  }
  return *reinterpret_cast<const real*>(reinterpret_cast<const char*>(xv.x) + (unsigned long long)c * xv.inc8);
}

// Where an entry of the operand vector comes from.  Normally x itself (xload); an epilogue that declares XFUSED = true supplies
// the entry through its own xg(col) instead -- the one-pass persistent CG (persistent.cuh) RECOMPUTES p_new[col] =
// r[col] - alpha Ap[col] + beta p[col] on the fly, which removes the separate vector-update phase and one of the two grid
// barriers of an iteration.
template<class Epi, class = void> struct epi_xfused { static constexpr bool value = false; };
template<class Epi> struct epi_xfused<Epi, typename std::enable_if<Epi::XFUSED>::type> { static constexpr bool value = true; };
template<bool SPLIT, class Epi>
__device__ __forceinline__ real xget(const Epi &epi, const XVec &xv, u32 c)
{
  if constexpr (epi_xfused<Epi>::value) return epi.xg(c);
  else return xload<SPLIT>(xv, c);
}

// In-row CSR accumulation step.  The reference host backend (host_based/sparse_matrix_operations.hpp:167-184), built with
// g++ -O3 for x86-64-v3, evaluates `dot += a*x` as a rounded multiply followed by a rounded add (GCC does not form FMA
// chains in reductions under generic tuning); the same two roundings are used here so that CSR results match it bit for bit.
__device__ __forceinline__ real madd(real a, real x, real acc) { return radd(acc, rmul(a, x)); }

#ifdef VCL_F32
// The FLOAT instantiation of the same reference loop is compiled differently by that build (pinned empirically,
// SURVEY 8c / DESIGN.md section 2): for unit-stride x GCC vectorises the gather loop 4 wide with an in-order reduction --
// rounded products, then adds -- and contracts the scalar remainder loop, i.e. the last (row length mod 4) entries, to
// fmaf; for strided x nothing is vectorised or contracted.  first_fused() = index of the first contracted entry of a row
// [s, e); madd_at() applies the matching operation to entry i.
__device__ __forceinline__ u32 first_fused(u32 s, u32 e, const XVec &xv) { return (xv.inc8 == (u32)sizeof(real)) ? s + ((e - s) & ~3u) : 0xffffffffu; }
__device__ __forceinline__ real madd_at(real a, real x, real acc, u32 i, u32 ff)
{
  const real q = fma(a, x, acc), p = radd(acc, rmul(a, x));
  return (i >= ff) ? q : p;
}
#else
__device__ __forceinline__ u32 first_fused(u32, u32, const XVec &) { return 0xffffffffu; }
__device__ __forceinline__ real madd_at(real a, real x, real acc, u32, u32) { return madd(a, x, acc); }
#endif


// The matrix arrays are read exactly once per product: they are streamed with an L2 evict-first policy so that they do not
// push the gathered x entries (re-used by the rows of the next planes, tens of MB of streamed matrix data later) out of L2.
__device__ __forceinline__ unsigned long long l2_evict_first_policy()
{
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
// A matrix that fits L2 together with the solver's vectors is re-read every iteration (BASELINE config 1: 67 MB of CSR arrays,
// 34 MB of vectors, 126 MB of L2): streaming it evict-first would throw it away each time.  mode: see CsrDev::l2_mode.
__device__ __forceinline__ unsigned long long l2_policy(int mode)
{
  unsigned long long pol;
  if (mode == 1)      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
  else if (mode == 2) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;\n" : "=l"(pol));
  else                asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
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
  real *y; int off, inc; real alpha, beta;
  typedef real Pre;
  bool sell_f32 = false;     // float SELL only: the reference build evaluates fma(alpha, dot, beta*y) there
  static constexpr int NQ = 0;
  static constexpr bool COO = false;
  __device__ __forceinline__ real init(real) const { return 0.0; }
  __device__ __forceinline__ real term_scale() const { return 1.0; }
  __device__ __forceinline__ bool skip() const { return false; }
  // pre(): the epilogue's own per-row operand, requested BEFORE the row's gather chain so that its latency overlaps
  __device__ __forceinline__ real pre(u32 r) const
  {
    return (beta != 0.0) ? y[(size_t)r * (size_t)inc + (size_t)off] : 0.0;
  }
  __device__ __forceinline__ void row(u32 r, real dot, real y_old)
  {
    size_t idx = (size_t)r * (size_t)inc + (size_t)off;
    // same operations as the reference host build: t = alpha*dot (rounded), then one fused beta*y + t
    if (beta != 0.0) y[idx] = sell_f32 ? fma(alpha, dot, rmul(beta, y_old)) : fma(beta, y_old, rmul(alpha, dot));
    else             y[idx] = rmul(alpha, dot);
  }
  __device__ __forceinline__ void finish(real *) {}
};

// coordinate_matrix product on the CSR index of the COO entries: y = (beta*y) + sum_k fma(alpha*a_k, x_k, .) in storage order
struct EpiCoo
{
  real *y; int off, inc; real alpha, beta;
  static constexpr int NQ = 0;
  static constexpr bool COO = true;
  typedef real Pre;
  __device__ __forceinline__ bool skip() const { return false; }
  __device__ __forceinline__ real pre(u32 r) const { return (beta != 0.0) ? y[(size_t)r * (size_t)inc + (size_t)off] : 0.0; }
  __device__ __forceinline__ real init(real y_old) const { return (beta != 0.0) ? rmul(y_old, beta) : 0.0; }
  __device__ __forceinline__ real term_scale() const { return alpha; }
  __device__ __forceinline__ void row(u32 r, real dot, real) { y[(size_t)r * (size_t)inc + (size_t)off] = dot; }
  __device__ __forceinline__ void finish(real *) {}
};

// ------------------------------------------------------------------------------------------------
// CSR, row-block streaming: persistent CTAs, two-deep TMA pipeline.
//
// Every CTA walks its row blocks bi = blockIdx.x, +gridDim.x, ...  While the 256 threads work on block j (one thread per
// row, entries read from shared memory), ONE elected thread has already issued two bulk copies (cp.async.bulk, the 1-D TMA
// path, L2 evict-first) that bring the value and column-index ranges of block j+1 into the other shared-memory buffer;
// completion is signalled through an mbarrier (complete_tx bytes), so no thread spends issue slots on the copy and the CTA
// always has a block in flight.  The block descriptors (row range, nnz range) are themselves software-pipelined in
// registers two and three blocks ahead, so that no dependent global load sits on the critical path.
//   class 0  normal block: 16-byte aligned enclosing range, <= CSR_CAP entries           -> TMA
//   class 1  one row longer than CSR_CAP: the whole CTA strides over it straight from global memory
//   class 2  range whose 16-byte padded end would pass the end of the arrays (last block) -> guarded synchronous staging
// Plans that break these limits (more than 256 rows in a block, or several rows with more than CSR_CAP entries -- e.g. the
// reference's own handle3() blocks, compressed_matrix.hpp:1152-1188: <= 1024 entries but any number of rows) never reach this
// kernel: the host checks every foreign plan once (vcl_plan_ok, common.cuh) and sends such products to csr_scalar_kernel.
// (Handling them HERE as a fourth block class cost the 256^3 product 8 %: 0.272 ms against 0.2515, profiles/ab_spmv_r2c.log.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar, unsigned long long pol)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n"
               :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

#ifndef CSR_NSTAGE
#define CSR_NSTAGE 2                                                               // shared-memory ring depth (blocks)
#endif
#define CSR_SMEM_BYTES (CSR_NSTAGE * CSR_STAGE * (int)(sizeof(real) + sizeof(u32)))  // dynamic shared memory
#ifndef CSR_MIN_CTAS
#define CSR_MIN_CTAS (CSR_NSTAGE <= 2 ? 4 : (CSR_NSTAGE == 3 ? 3 : 2))
#endif
// Resident CTAs per SM the stand-alone product kernels are compiled for.  The float build stages 8 instead of 12 bytes per entry
// (32.8 KB per CTA), so more CTAs fit one SM, and its kernels are bound by the latency of the gather chain, not by HBM: measured at
// 256^3 (profiles/ab_float_ctas_r2n.log) CSR 0.1967 / 0.1898 / 0.2112 ms and SELL-32 0.1969 / 0.1754 / 0.1680 ms with 4 / 5 / 6 CTAs
// (48 / 40 registers, no spills); with the raw descriptor registers of the current SELL kernel 5 CTAs are ahead of 6 (0.1747 / 0.1775 ms).
// The persistent cooperative kernels keep CSR_MIN_CTAS (they spill below 64 registers).
#if defined(VCL_F32) && CSR_NSTAGE <= 2
#define CSR_STREAM_MIN_CTAS 5
#define SELL_MIN_CTAS(NQ, CT) 5
#else
#define CSR_STREAM_MIN_CTAS CSR_MIN_CTAS
#define SELL_MIN_CTAS(NQ, CT) CSR_MIN_CTAS
#endif

struct CsrBlockDesc { u32 r0, r1, n0, n1; };

__device__ __forceinline__ int csr_block_class(const CsrBlockDesc &d, u32 nnz)
{
  if (d.n1 - d.n0 > CSR_CAP) return 1;
  const u32 a0 = d.n0 & ~3u;
  const u32 cnt4 = (d.n1 - a0 + 3u) & ~3u;
  return (cnt4 == 0u || a0 + cnt4 > nnz) ? 2 : 0;
}

// One row, entries [j, e) of the staged block, sequential accumulation chain in storage order.  All gathers of up to 8
// entries are issued before the first multiply-add (predicated, so a 7-point row costs ONE round trip to L2, not three).
// COO == true: coordinate_matrix semantics (host_based/sparse_matrix_operations.hpp:1233-1246): the chain starts at beta*y
// and every term is fma(alpha*a, x, .) -- what the reference build does for that format.
template<bool SPLIT, bool COO, class Epi>
__device__ __forceinline__ real csr_row_dot(const Epi &epi, const real *s_val, const u32 *s_col, u32 j, u32 e, const XVec &xv,
                                              real init = 0.0, real alpha = 1.0)
{
  real dot = COO ? init : 0.0;
  const u32 ff = first_fused(j, e, xv);
  for (; j < e; j += 8)
  {
    real v[8], xx[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
      const bool ok = j + k < e;
      xx[k] = ok ? xget<SPLIT>(epi, xv, s_col[j + k]) : 0.0;
      v[k] = ok ? s_val[j + k] : 0.0;
    }
    // empty slots hold v = x = +0.0: adding +0.0 never changes the bits of `dot` (a round-to-nearest sum that starts at
    // +0.0 cannot be -0.0), so the chain needs no predicates and stays bit-identical to the sequential reference
#pragma unroll
    for (int k = 0; k < 8; ++k) dot = COO ? fma(rmul(alpha, v[k]), xx[k], dot) : madd_at(v[k], xx[k], dot, j + k, ff);
  }
  return dot;
}

// The kernel body is a device function so that the persistent solver kernels (persistent.cuh) can run the same product
// between two grid-wide barriers; REPEATED: the mbarriers are invalidated on exit because the body will run again.
// REPEATED (persistent kernels): the matrix does not change between calls, so (i) the block descriptors of the prologue are
// computed once and parked in shared memory, (ii) the mbarriers live on (their phase bits travel in `carry`), and (iii) the
// copy of this CTA's first block for the NEXT call is issued at the end of the current one -- it is in flight while the
// kernel does its vector update and waits at the grid barrier.
struct CsrCarry { unsigned phase; int primed; };

// SPLIT: row-partitioned launch (halo push at the head, flag wait before the boundary blocks).  XS: x is addressed as [owned | halo] in
// two buffers (xload<true>); XS = false with SPLIT = true is the solver form in which the gathered vector and its halo are ONE
// contiguous array inside the peer window (dist.cu: gather vector), so every block uses the plain one-base addressing.
template<class Epi, bool SPLIT, bool REPEATED, bool XS = SPLIT>
__device__ __forceinline__ void csr_stream_body(const CsrDev &A, const XVec &xv, Epi &epi, CsrCarry *carry = nullptr, bool drain = false)
{
  constexpr int S = CSR_NSTAGE;
  __shared__ CsrBlockDesc s_keepD[REPEATED ? S : 1];
  __shared__ u32 s_keepR[2];
  __shared__ u32 s_keepMy[REPEATED ? 2 * CSR_BLOCK_THREADS : 1];
  extern __shared__ __align__(128) unsigned char csr_smem[];
  __shared__ __align__(8) unsigned long long s_bar[S];
  __shared__ real s_red[(Epi::NQ > 0 ? Epi::NQ : 1) * 32];
  real *s_val0 = reinterpret_cast<real*>(csr_smem);
  u32 *s_col0 = reinterpret_cast<u32*>(csr_smem + S * CSR_STAGE * sizeof(real));

  if (epi.skip()) return;
  const int tid = threadIdx.x;
  const int step = (int)gridDim.x;
  const unsigned long long pol = l2_policy(A.l2_mode);
#ifdef VCL_PEER_DEBUG
  if (SPLIT && A.dbg && blockIdx.x == 0 && tid == 0) A.dbg[(A.dbg_seq % 1024) * 4 + 2] = global_ns();
#endif
  bool halo_ready = !(SPLIT && A.wait_mask != 0u);

  const bool resume = REPEATED && carry->primed != 0;     // uniform over the CTA
  if (REPEATED && drain)
  {
    // last call of a persistent kernel: a CTA must not exit while a bulk copy into its shared memory is in flight
    if (resume)
    {
#pragma unroll
      for (int i = 0; i < S - 1; ++i)
        if ((int)blockIdx.x + i * (int)gridDim.x < A.nblk && csr_block_class(s_keepD[i], A.nnz) == 0) mbar_wait(&s_bar[i], (carry->phase >> i) & 1u);
    }
    return;
  }
  if (!resume)
  {
    if (tid == 0)
    {
#pragma unroll
      for (int i = 0; i < S; ++i) mbar_init(&s_bar[i], 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
  }

  // elected thread: bring block `d` into buffer `buf`
  auto issue = [&](const CsrBlockDesc &d, int buf)
  {
    const u32 a0 = d.n0 & ~3u;
    const u32 cnt4 = (d.n1 - a0 + 3u) & ~3u;
    fence_proxy_async();                                   // earlier generic-proxy accesses to this buffer are done (barrier)
    mbar_expect_tx(&s_bar[buf], cnt4 * (unsigned)(sizeof(real) + sizeof(u32)));
    tma_load_1d(s_val0 + buf * CSR_STAGE, A.va + a0, cnt4 * (unsigned)sizeof(real), &s_bar[buf], pol);
    tma_load_1d(s_col0 + buf * CSR_STAGE, A.ci + a0, cnt4 * 4u, &s_bar[buf], pol);
  };

  // ---- prologue: D[i] = descriptor of this CTA's i-th block, (r0_S, r1_S) = row range of the S-th ----
  int bi = blockIdx.x;
  CsrBlockDesc D[S];
  u32 r0_S = 0, r1_S = 0;
  u32 my_s = 0, my_e = 0;                                  // this thread's row of the current block
  unsigned phase = 0;                                      // bit b: parity to wait for on buffer b
  if (resume)
  {
    // descriptors from the first call; the copy of the first block(s) was issued when the previous call ended
#pragma unroll
    for (int i = 0; i < S; ++i) D[i] = s_keepD[i];
    r0_S = s_keepR[0]; r1_S = s_keepR[1];
    my_s = s_keepMy[tid]; my_e = s_keepMy[CSR_BLOCK_THREADS + tid];
    phase = carry->phase;
  }
  else
  {
#pragma unroll
    for (int i = 0; i < S; ++i)
    {
      D[i].r0 = D[i].r1 = D[i].n0 = D[i].n1 = 0;
      if (bi + i * step < A.nblk) { D[i].r0 = A.blk_start[bi + i * step]; D[i].r1 = A.blk_end[bi + i * step]; }
    }
    if (bi + S * step < A.nblk) { r0_S = A.blk_start[bi + S * step]; r1_S = A.blk_end[bi + S * step]; }
#pragma unroll
    for (int i = 0; i < S; ++i)
      if (bi + i * step < A.nblk) { D[i].n0 = A.rp[D[i].r0]; D[i].n1 = A.rp[D[i].r1]; }
    if (bi < A.nblk && (u32)tid < D[0].r1 - D[0].r0) { my_s = A.rp[D[0].r0 + tid]; my_e = A.rp[D[0].r0 + tid + 1]; }
    if (REPEATED)
    {
      if (tid == 0)
      {
#pragma unroll
        for (int i = 0; i < S; ++i) s_keepD[i] = D[i];
        s_keepR[0] = r0_S; s_keepR[1] = r1_S;
      }
      s_keepMy[tid] = my_s; s_keepMy[CSR_BLOCK_THREADS + tid] = my_e;
    }
    if (tid == 0)
    {
#pragma unroll
      for (int i = 0; i < S - 1; ++i)
        if (bi + i * step < A.nblk && csr_block_class(D[i], A.nnz) == 0) issue(D[i], i);
    }
  }
  int buf = 0;                                             // buffer of the current block

  if (SPLIT && A.push != nullptr)
  {
    // the copy of this CTA's first row block is already in flight; now send x's boundary entries (remote stores over NVLink)
    halo_push_share(*A.push, A.push_idx, xv.x, A.wait_seq, A.push_ticket);
  }

  for (; bi < A.nblk; bi += step)
  {
    // ---- keep the pipeline full: copy of block j+S-1, row pointers of j+1, nnz range of j+S, row range of j+S+1 ----
    const int buf_issue = (buf == 0) ? S - 1 : buf - 1;    // == (buf + S - 1) % S: the buffer block j-1 has just released
    if (tid == 0 && bi + (S - 1) * step < A.nblk && csr_block_class(D[S - 1], A.nnz) == 0) issue(D[S - 1], buf_issue);
    u32 nx_s = 0, nx_e = 0, n0_S = 0, n1_S = 0, r0_T = 0, r1_T = 0;
    if (bi + step < A.nblk && (u32)tid < D[1].r1 - D[1].r0) { nx_s = A.rp[D[1].r0 + tid]; nx_e = A.rp[D[1].r0 + tid + 1]; }
    if (bi + S * step < A.nblk) { n0_S = A.rp[r0_S]; n1_S = A.rp[r1_S]; }
    if (bi + (S + 1) * step < A.nblk) { r0_T = A.blk_start[bi + (S + 1) * step]; r1_T = A.blk_end[bi + (S + 1) * step]; }

    if (SPLIT && !halo_ready && bi >= A.wait_from)
    {
      // boundary row blocks come last in the list: by now the neighbours' pushes have normally landed
#ifdef VCL_PEER_DEBUG
      const u64 t_w = global_ns();
#endif
      if (tid < VCL_MAX_PEERS && ((A.wait_mask >> tid) & 1u)) peer_wait(A.wait_flags + tid, A.wait_seq, A.err, 1, tid);
      __syncthreads();
#ifdef VCL_PEER_DEBUG
      if (A.dbg && tid == 0) atomicMax(A.dbg + (A.dbg_seq % 1024) * 4 + 3, global_ns() - t_w);
#endif
      halo_ready = true;
    }

    const CsrBlockDesc cur = D[0];
    const u32 nrows = cur.r1 - cur.r0;
    const int cls = csr_block_class(cur, A.nnz);
    if (cls == 1)
    {
      // one long row: the whole CTA strides over it (summation order differs from the sequential reference; tolerance-level parity)
      real part[1] = {0.0};
      for (u32 k = cur.n0 + tid; k < cur.n1; k += CSR_BLOCK_THREADS)
        part[0] = fma(A.va[k], xget<XS>(epi, xv, A.ci[k]), part[0]);
      __shared__ real s_long[32];
      block_sum<1>(part, s_long);
      if (tid == 0)
      {
        const typename Epi::Pre pre = epi.pre(cur.r0);
        epi.row(cur.r0, Epi::COO ? fma(epi.term_scale(), part[0], epi.init(pre)) : part[0], pre);
      }
    }
    else
    {
      typename Epi::Pre pre = typename Epi::Pre();
      if ((u32)tid < nrows) pre = epi.pre(cur.r0 + tid);
      const u32 a0 = cur.n0 & ~3u;
      const real *s_val = s_val0 + buf * CSR_STAGE;
      const u32 *s_col = s_col0 + buf * CSR_STAGE;
      if (cls == 0)
      {
        mbar_wait(&s_bar[buf], (phase >> buf) & 1u);
        phase ^= 1u << buf;
      }
      else
      {
        // guarded staging of the last few entries of the arrays (never reads past nnz)
        real *wv = s_val0 + buf * CSR_STAGE; u32 *wc = s_col0 + buf * CSR_STAGE;
        for (u32 i = tid; a0 + i < cur.n1; i += CSR_BLOCK_THREADS) { wv[i] = A.va[a0 + i]; wc[i] = A.ci[a0 + i]; }
        __syncthreads();
      }
      // Rows of up to CSR_LONG_ROW entries: one thread per row, sequential chain (bit-identical to the reference's loop).
      // Longer rows would leave that thread gathering alone for dozens of L2 round trips while the CTA waits: each warp
      // takes the long rows of its own 32 rows one after the other, all lanes striding over the staged entries, and sums
      // with shuffles (no block-level synchronisation; summation order differs from the reference: tolerance-level parity,
      // like the whole-CTA path for rows beyond CSR_CAP).  COO keeps the sequential chain (its alpha/beta form is defined by it).
      const bool is_long = !Epi::COO && (my_e - my_s) > (u32)CSR_LONG_ROW;
      if ((u32)tid < nrows && !is_long)
      {
        // row-partitioned slabs: INTERIOR blocks (bi < wait_from) reference owned columns only, so they gather with the plain
        // one-base addressing; only the boundary blocks (3 % at 256^3 per GPU) pay the base selection of the [owned | halo] form
        // (measured: 0.2615 -> 0.2489 ms per product at world 1, 0.2546 -> 0.2507 ms on 2 GPUs).
        // Only for the plain product (Epi::NQ == 0): with the fused solver epilogues the second copy of the row loop costs more
        // than the selection saves (fused 512^3 CG product on 2 GPUs 284 -> 298 us, profiles/ab_interior_r2r.log).
        constexpr bool DUAL = XS && Epi::NQ == 0;
        const real dot = (!DUAL || bi >= A.wait_from)
                             ? csr_row_dot<XS, Epi::COO>(epi, s_val, s_col, my_s - a0, my_e - a0, xv, epi.init(pre), epi.term_scale())
                             : csr_row_dot<false, Epi::COO>(epi, s_val, s_col, my_s - a0, my_e - a0, xv, epi.init(pre), epi.term_scale());
        epi.row(cur.r0 + tid, dot, pre);
      }
      unsigned longs = __ballot_sync(0xffffffffu, is_long);
      while (longs)
      {
        const int src = __ffs(longs) - 1;
        longs &= longs - 1u;
        const u32 ls = __shfl_sync(0xffffffffu, my_s, src) - a0, le = __shfl_sync(0xffffffffu, my_e, src) - a0;
        real part = 0.0;
        for (u32 k = ls + (u32)(tid & 31); k < le; k += 32u) part = fma(s_val[k], xget<XS>(epi, xv, s_col[k]), part);
        part = warp_sum(part);
        if ((tid & 31) == src) epi.row(cur.r0 + tid, part, pre);
      }
    }
    __syncthreads();                                       // buffer `buf` may be refilled from the next iteration on
    // ---- rotate the descriptor pipeline ----
#pragma unroll
    for (int i = 0; i + 1 < S; ++i) D[i] = D[i + 1];
    D[S - 1].r0 = r0_S; D[S - 1].r1 = r1_S; D[S - 1].n0 = n0_S; D[S - 1].n1 = n1_S;
    r0_S = r0_T; r1_S = r1_T;
    my_s = nx_s; my_e = nx_e;
    buf = (buf + 1 == S) ? 0 : buf + 1;
  }
  epi.finish(s_red);
  if (REPEATED)
  {
    __syncthreads();                                       // every buffer is free, s_keep* is complete
    if (tid == 0)
    {
#pragma unroll
      for (int i = 0; i < S - 1; ++i)
        if ((int)blockIdx.x + i * step < A.nblk && csr_block_class(s_keepD[i], A.nnz) == 0) issue(s_keepD[i], i);
    }
    carry->phase = phase;
    carry->primed = 1;
  }
}

template<class Epi, bool SPLIT, bool XS = SPLIT>
__global__ void __launch_bounds__(CSR_BLOCK_THREADS, CSR_STREAM_MIN_CTAS)
csr_stream_kernel(CsrDev A, XVec xv, Epi epi)
{
  csr_stream_body<Epi, SPLIT, false, XS>(A, xv, epi);
}


// Plan-free CSR kernel: one thread per row, sequential order (used when no row blocks are supplied or the
// arrays are not 16-byte aligned, e.g. oddly offset user-wrapped buffers).
template<class Epi>
__global__ void __launch_bounds__(256)
csr_scalar_kernel(CsrDev A, XVec xv, Epi epi)
{
  __shared__ real s_red[(Epi::NQ > 0 ? Epi::NQ : 1) * 32];
  if (epi.skip()) return;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < A.rows; r += (long long)gridDim.x * blockDim.x)
  {
    const typename Epi::Pre pre = epi.pre((u32)r);
    real dot = Epi::COO ? epi.init(pre) : 0.0;
    const u32 e = A.rp[r + 1];
    const u32 ff = first_fused(A.rp[r], e, xv);
    for (u32 k = A.rp[r]; k < e; ++k)
      dot = Epi::COO ? fma(rmul(epi.term_scale(), A.va[k]), xload<false>(xv, A.ci[k]), dot) : madd_at(A.va[k], xload<false>(xv, A.ci[k]), dot, k, ff);
    epi.row((u32)r, dot, pre);
  }
  epi.finish(s_red);
}

// ------------------------------------------------------------------------------------------------
// SELL-C-sigma (sigma = 1).  Slices are stored back to back, so the slices of one CTA pass (256/C of them: 8 for the default
// C = 32) form ONE contiguous, 16-byte aligned range of values and of column indices: it is brought into shared memory
// by the same two-stage TMA pipeline as a CSR row block (block j+1 is in flight while block j is computed), then one
// thread per row walks its slice-column-major entries (stride C in shared memory: consecutive rows hit consecutive banks)
// with one fma chain, all gathers of up to 8 entries issued before the first fma.  Zero-valued (padding) slots never
// touch x (cuda/sparse_matrix_operations.hpp:2231, host :1833; gathering first and masking afterwards -- so that the x load does
// not wait for the value -- was measured: 0.2667 -> 0.2685 ms at 256^3, no gain, profiles/ab_sell_r2j.log).  Passes whose range does not fit the staging buffer, or
// C not a multiple of 4 / larger than the CTA, take the direct path.
// The multiply-adds are fused: that is what the reference host build does for SELL (oracle/vcl_oracle.c, ARITHMETIC).
// ------------------------------------------------------------------------------------------------
// one SELL row from the staged slice-column-major entries: entry j of the row sits at idx + j*C; 8 gathers in flight, one fma chain
template<bool SPLITV>
__device__ __forceinline__ real sell_row_acc(const real *s_val, const u32 *s_col, u32 idx, u32 w, u32 C, const XVec &xv)
{
  real acc = 0.0;
  for (u32 j = 0; j < w; j += 8, idx += 8 * C)
  {
    real v[8], xx[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
      v[k] = (j + k < w) ? s_val[idx + k * C] : 0.0;
      xx[k] = nonzero(v[k]) ? xload<SPLITV>(xv, s_col[idx + k * C]) : 0.0;
    }
    // zero (padding / empty) slots have v = x = +0.0 and leave the bits of `acc` unchanged: no predicates needed
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fma(xx[k], v[k], acc);
  }
  return acc;
}

// CT: slice height known at compile time (32, the reference's default, sliced_ell_matrix.hpp:146-147) or 0 = read A.C.  With a
// constant C the in-slice strides fold into the immediate offsets of the shared-memory loads and tid / C, tid % C become a
// shift and a mask; the generic form spent 31 % of its issue slots on IMAD address arithmetic and two 32-bit divisions per
// pass (ncu source page, profiles/ncu_summary_r2.md) and ran 8 % behind the CSR kernel although it moves 3.5 % fewer bytes.
template<class Epi, bool PERM, int CT, bool SPLIT = false>      // PERM: SELL-C-sigma (storage row -> matrix row through A.perm); false: the reference's layout
__global__ void __launch_bounds__(CSR_BLOCK_THREADS, SELL_MIN_CTAS(Epi::NQ, CT))
sell_kernel(SellDev A, XVec xv, Epi epi)
{
  constexpr int S = 2;                                     // this kernel is written for a two-stage ring (it uses the first two CSR stages)
  extern __shared__ __align__(128) unsigned char csr_smem[];
  __shared__ __align__(8) unsigned long long s_bar[S];
  __shared__ real s_red[(Epi::NQ > 0 ? Epi::NQ : 1) * 32];
  real *s_val0 = reinterpret_cast<real*>(csr_smem);
  u32 *s_col0 = reinterpret_cast<u32*>(csr_smem + S * CSR_STAGE * sizeof(real));
  static_assert(CSR_NSTAGE >= 2, "sell_kernel needs two stages of the CSR ring");

  if (epi.skip()) return;
  const real * __restrict__ va = A.va;
  const u32 * __restrict__ ci = A.ci;
  const int tid = threadIdx.x;
  const int step = (int)gridDim.x;
  const unsigned long long pol = l2_evict_first_policy();
  const u32 C = CT ? (u32)CT : (u32)A.C;
  const u32 nslices = ((u32)A.rows - 1u) / C + 1u;
  const u32 t_slice = (u32)tid / C, t_lane = (u32)tid % C;                 // this thread's slice within a pass and row within the slice
  const u32 spb = C <= CSR_BLOCK_THREADS ? CSR_BLOCK_THREADS / C : 1u;      // slices per CTA pass
  const int nblocks = (int)((nslices + spb - 1) / spb);
  const bool can_stage = (C % 4u) == 0u && C <= CSR_BLOCK_THREADS &&
                         ((reinterpret_cast<uintptr_t>(va) | reinterpret_cast<uintptr_t>(ci)) & 15u) == 0u;

  if (tid == 0)
  {
    mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  // Row-partitioned slabs: the passes that gather from the halo are the first and the last ones of the slab (a 1-D row partition);
  // every CTA walks the passes rotated by half the list, so that those come in the MIDDLE of its sequence -- by then the neighbours'
  // entries have arrived -- instead of stalling the CTAs whose first pass lies in the first plane.  phys(): logical -> stored pass.
  const int rot = SPLIT ? nblocks / 2 : 0;
  auto phys = [&](int b) { const int q = b + rot; return q >= nblocks ? q - nblocks : q; };
  // Descriptors of a pass are fetched one / two passes ahead and kept RAW in registers: any arithmetic on a freshly loaded value would
  // stall the warp at that instruction until the load returns (in-order issue) -- the first version computed `end` and `first`
  // right after the loads and spent 8.7 % of its stall samples there (ncu source page, profiles/ncu_summary_r2h.md).
  // element range of pass b: base = r.x, end = r.y + r.z * C
  auto range_raw = [&](int b, u32 &r0, u32 &r1, u32 &r2)
  {
    const u32 s0 = (u32)phys(b) * spb, s1 = min(s0 + spb, nslices);
    r0 = A.bs[s0]; r1 = A.bs[s1 - 1]; r2 = A.cpb[s1 - 1];
  };
  auto staged = [&](u32 base, u32 end) { return can_stage && end > base && end - base <= CSR_CAP; };
  auto issue = [&](u32 base, u32 end, int buf)
  {
    const u32 cnt = end - base;                            // multiple of C, hence of 4; base likewise
    fence_proxy_async();
    mbar_expect_tx(&s_bar[buf], cnt * (unsigned)(sizeof(real) + sizeof(u32)));
    tma_load_1d(s_val0 + buf * CSR_STAGE, va + base, cnt * (unsigned)sizeof(real), &s_bar[buf], pol);
    tma_load_1d(s_col0 + buf * CSR_STAGE, ci + base, cnt * 4u, &s_bar[buf], pol);
  };
  // this thread's row of pass b: slice width and (raw) start of its slice; first entry = start + t_lane
  auto my_row_raw = [&](int b, u32 &w, u32 &start)
  {
    const u32 slice = (u32)phys(b) * spb + t_slice;
    w = 0; start = 0;
    if ((u32)tid < spb * C && slice < nslices) { w = A.cpb[slice]; start = A.bs[slice]; }
  };
  int b = blockIdx.x;
  u32 base_c = 0, end_c = 0, base_n = 0, end_n = 0, w_c = 0, first_c = 0;
  if (b < nblocks)
  {
    u32 q0, q1, q2, st0;
    range_raw(b, q0, q1, q2); base_c = q0; end_c = q1 + q2 * C;
    my_row_raw(b, w_c, st0); first_c = st0 + t_lane;
  }
  if (b + step < nblocks) { u32 q0, q1, q2; range_raw(b + step, q0, q1, q2); base_n = q0; end_n = q1 + q2 * C; }
  if (tid == 0 && b < nblocks && staged(base_c, end_c)) issue(base_c, end_c, 0);
  unsigned phase = 0;
  int buf = 0;
  bool halo_ready = !(SPLIT && A.wait_mask != 0u);
  // "does pass b gather from the halo?" is fetched one pass ahead like the other descriptors (a load at the top of the pass would
  // sit on the critical path of every pass: measured +19 us per product at 256^3)
  unsigned char need_c = (SPLIT && A.needs != nullptr && b < nblocks) ? A.needs[phys(b)] : (unsigned char)0;
  if (SPLIT && A.push != nullptr) halo_push_share(*A.push, A.push_idx, xv.x, A.wait_seq, A.push_ticket);   // the first copy is already in flight

  for (; b < nblocks; b += step, buf ^= 1)
  {
    const unsigned char need_n = (SPLIT && A.needs != nullptr && b + step < nblocks) ? A.needs[phys(b + step)] : (unsigned char)0;
    if (SPLIT && !halo_ready && need_c)
    {
      // first pass of this CTA that gathers from the halo: by now the neighbours' pushes have normally landed
      if (tid < VCL_MAX_PEERS && ((A.wait_mask >> tid) & 1u)) peer_wait(A.wait_flags + tid, A.wait_seq, A.err, 1, tid);
      __syncthreads();
      halo_ready = true;
    }
    if (tid == 0 && b + step < nblocks && staged(base_n, end_n)) issue(base_n, end_n, buf ^ 1);
    u32 raw0 = 0, raw1 = 0, raw2 = 0, w_n = 0, start_n = 0;     // consumed only at the end of this pass
    if (b + 2 * step < nblocks) range_raw(b + 2 * step, raw0, raw1, raw2);
    if (b + step < nblocks) my_row_raw(b + step, w_n, start_n);

    const u32 s0 = (u32)phys(b) * spb, s1 = min(s0 + spb, nslices);
    if (staged(base_c, end_c))
    {
      // (s1 - s0) * C <= 256 here: one row per thread
      long long r = (long long)(s0 + t_slice) * C + t_lane;                          // storage row ...
      if (PERM && (u32)tid < (s1 - s0) * C) r = (long long)A.perm[r];  // ... -> matrix row (padding: 0xFFFFFFFF)
      const bool active = (u32)tid < (s1 - s0) * C && r < A.rows;
      typename Epi::Pre pre = typename Epi::Pre();
      if (active) pre = epi.pre((u32)r);
      const real *s_val = s_val0 + buf * CSR_STAGE;
      const u32 *s_col = s_col0 + buf * CSR_STAGE;
      mbar_wait(&s_bar[buf], (phase >> buf) & 1u);
      phase ^= 1u << buf;
      if (active)
      {
        // passes without halo columns gather with the plain one-base addressing (see csr_stream_body)
        constexpr bool DUAL = SPLIT && Epi::NQ == 0;       // plain product only, as in csr_stream_body
        const real acc = (SPLIT && (!DUAL || need_c)) ? sell_row_acc<true>(s_val, s_col, first_c - base_c, w_c, C, xv)
                                                      : sell_row_acc<false>(s_val, s_col, first_c - base_c, w_c, C, xv);
        epi.row((u32)r, acc, pre);
      }
    }
    else
    {
      // direct path (very wide slices, C not a multiple of 4, or C > 256): coalesced 8-/4-byte loads straight from global
      for (u32 t = tid; t < (s1 - s0) * C; t += CSR_BLOCK_THREADS)
      {
        const u32 slice = s0 + t / C;
        long long r = (long long)slice * C + (t % C);
        if (PERM) r = (long long)A.perm[r];
        if (r >= A.rows) continue;
        const u32 w = A.cpb[slice];
        size_t idx = (size_t)A.bs[slice] + (t % C);
        real acc = 0.0;
        u32 j = 0;
        for (; j + 4 <= w; j += 4, idx += 4 * (size_t)C)
        {
          const real v0 = va[idx], v1 = va[idx + C], v2 = va[idx + 2 * (size_t)C], v3 = va[idx + 3 * (size_t)C];
          const u32 c0 = ci[idx], c1 = ci[idx + C], c2 = ci[idx + 2 * (size_t)C], c3 = ci[idx + 3 * (size_t)C];
          const real x0 = (v0 != 0.0) ? xload<SPLIT>(xv, c0) : 0.0;
          const real x1 = (v1 != 0.0) ? xload<SPLIT>(xv, c1) : 0.0;
          const real x2 = (v2 != 0.0) ? xload<SPLIT>(xv, c2) : 0.0;
          const real x3 = (v3 != 0.0) ? xload<SPLIT>(xv, c3) : 0.0;
          if (v0 != 0.0) acc = fma(x0, v0, acc);
          if (v1 != 0.0) acc = fma(x1, v1, acc);
          if (v2 != 0.0) acc = fma(x2, v2, acc);
          if (v3 != 0.0) acc = fma(x3, v3, acc);
        }
        for (; j < w; ++j, idx += C)
        {
          const real v0 = va[idx];
          if (v0 != 0.0) acc = fma(xload<SPLIT>(xv, ci[idx]), v0, acc);
        }
        epi.row((u32)r, acc, epi.pre((u32)r));
      }
    }
    __syncthreads();                                       // buffer `buf` may be refilled from the next iteration on
    base_c = base_n; end_c = end_n; base_n = raw0; end_n = raw1 + raw2 * C; w_c = w_n; first_c = start_n + t_lane;
    need_c = need_n;
  }
  epi.finish(s_red);
}

// ------------------------------------------------------------------------------------------------
// ELL / HYB (cuda/sparse_matrix_operations.hpp:1747-1838, :2298-2400).  The ELL arrays are column-major over the rows, so
// a warp that owns 32 consecutive rows reads 256 / 128 contiguous bytes per slot: the loads are coalesced as they are and
// need no staging; they are issued 8 slots at a time (values and column indices together, L2 evict-first, no L1
// allocation), then the gathers of the non-zero slots, then one fma chain -- the reference host build fuses the ELL
// update (oracle/vcl_oracle.c, ARITHMETIC).  HYB: the same thread then walks its CSR tail with the unfused multiply-add
// the reference build uses there.  One thread per row, persistent grid-stride over rows.
// ------------------------------------------------------------------------------------------------
struct EllDev
{
  int rows, internal_rows, width;
  const u32 *coords; const real *elements;
  const u32 *csr_rows, *csr_cols; const real *csr_elements;      // HYB tail, NULL for plain ELL
};

__device__ __forceinline__ real ldg_stream(const real *p, unsigned long long pol)
{
  real v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint." VCL_PTX_REAL " %0, [%1], %2;\n" : "=" VCL_PTX_REG(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ u32 ldg_stream(const u32 *p, unsigned long long pol)
{
  u32 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;\n" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}

#ifndef ELL_ILP
#ifdef VCL_F32
#define ELL_ILP 8            // float: half the bytes per slot and half the registers -- twice the slots in flight
#else
#define ELL_ILP 4
#endif
#endif
#ifndef ELL_FUSED_CTAS
#define ELL_FUSED_CTAS 4      // fused solver epilogues (3 accumulators + prefetched operands) spill under the 42-register cap of 6 CTAs/SM
#endif
template<class Epi>
__global__ void __launch_bounds__(CSR_BLOCK_THREADS, (Epi::NQ > 0 ? ELL_FUSED_CTAS : 6))
ell_kernel(EllDev A, XVec xv, Epi epi)
{
  __shared__ real s_red[(Epi::NQ > 0 ? Epi::NQ : 1) * 32];
  if (epi.skip()) return;
  const unsigned long long pol = l2_evict_first_policy();
  const size_t IR = (size_t)A.internal_rows;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < A.rows; r += (long long)gridDim.x * blockDim.x)
  {
    const typename Epi::Pre pre = epi.pre((u32)r);
    u32 t0 = 0, t1 = 0;
    if (A.csr_rows) { t0 = A.csr_rows[r]; t1 = A.csr_rows[r + 1]; }
    real acc = 0.0;
    const real *pv = A.elements + r;
    const u32 *pc = A.coords + r;
    for (int j = 0; j < A.width; j += ELL_ILP, pv += ELL_ILP * IR, pc += ELL_ILP * IR)
    {
      real v[ELL_ILP], xx[ELL_ILP]; u32 c[ELL_ILP];
#pragma unroll
      for (int k = 0; k < ELL_ILP; ++k)
      {
        const bool ok = j + k < A.width;
        v[k] = ok ? ldg_stream(pv + k * IR, pol) : 0.0;
        c[k] = ok ? ldg_stream(pc + k * IR, pol) : 0u;
      }
#pragma unroll
      for (int k = 0; k < ELL_ILP; ++k) xx[k] = nonzero(v[k]) ? xload<false>(xv, c[k]) : 0.0;
#pragma unroll
      for (int k = 0; k < ELL_ILP; ++k) acc = fma(xx[k], v[k], acc);      // zero slots: +0.0, bits unchanged
    }
    const u32 ff = first_fused(t0, t1, xv);     // float build of the reference: same vectorised loop shape as the CSR product
    for (u32 k = t0; k < t1; ++k) acc = madd_at(A.csr_elements[k], xload<false>(xv, A.csr_cols[k]), acc, k, ff);
    epi.row((u32)r, acc, pre);
  }
  epi.finish(s_red);
}
} // namespace VCL_NS
