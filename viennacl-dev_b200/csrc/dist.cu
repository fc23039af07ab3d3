// dist.cu -- row-partitioned SpMV (CSR or SELL slabs) and Krylov solvers (CG, CG + diagonal preconditioner, BiCGStab, GMRES) across the
// GPUs of one box, one process per GPU.
// No counterpart in the reference ("Partition of data is left to the user", doc/manual/multi-device.dox:9).
//
// Rank g owns a contiguous block of rows.  At create time the locally stored (global) column indices are analysed on the
// device: columns outside the owned range become the halo; they are renumbered to [owned | halo] (halo sorted by global id,
// hence grouped by owner), every owner learns which of its entries each peer needs (send lists), and the CSR row blocks
// are split into INTERIOR blocks (no halo column) and BOUNDARY blocks.
// Transport "p2p" (default; peer.cuh): every rank maps a window of every other rank's memory (CUDA IPC over NVLink).
//   One product    = ONE csr_stream_kernel / sell_kernel launch: its first CTAs push this rank's boundary entries to the neighbours'
//                    receive buffers (+ flag), interior row blocks come first, boundary blocks wait for the neighbours' flags.
//   One CG iteration = cg_update_kernel + the fused SpMV kernel whose last CTA all-reduces {<r,r>, <Ap,Ap>, <p,Ap>} through the
//                    windows and advances alpha/beta/convergence -- 2 launches, one stream, no communication kernels, no host
//                    round trip.  (halo_push_kernel, the round-1 stand-alone push, survives behind VCL_B200_SEPARATE_PUSH.)
// Transport "nccl" (VCL_B200_DIST_TRANSPORT=nccl, or when IPC mapping is not possible): pack -> ncclSend/ncclRecv on the
//   communication stream || interior blocks; boundary blocks after the halo event; ncclAllReduce of 3 doubles; a
//   one-thread kernel advances the scalars.
// Both sum the rank-local totals so that every rank takes the same decisions (Chronopoulos/Gear CG needs a single
// reduction per iteration, cg.hpp:116-118).
#include "fused_kernels.cuh"
#include "launch.cuh"
#include "nvtx.cuh"
#include "gmres_host.cuh"
#include "gmres_launch.cuh"
#include "blas1.cuh"
#include "nccl_dyn.cuh"
#include <cmath>
#include <algorithm>

#ifdef VCL_F32
#error "the row-partitioned path is built in double precision only"
#endif
using namespace vcl_f64;

struct ViennaCLB200DistCsr_impl
{
  long long global_rows = 0, rb = 0, re = 0;
  int n = 0, nnz = 0;
  const u32 *rp = nullptr, *ci_global = nullptr; const double *va = nullptr;     // caller-owned
  u32 *ci_local = nullptr;                                                        // remapped copy (owned here)
  int n_halo = 0;
  std::vector<int> recv_cnt, recv_off, send_cnt, send_off;
  int total_send = 0;
  u32 *send_idx = nullptr;           // local indices to pack, grouped by destination
  double *send_buf = nullptr, *halo_buf = nullptr;
  u32 *blk = nullptr; int nblk = 0;
  int n_interior = 0, n_boundary = 0;
  double *tmp_sums = nullptr;        // 3 doubles: interior totals
  cudaEvent_t ev_x = nullptr, ev_halo = nullptr;
  // ---- peer-memory transport ----
  bool p2p = false;
  void *win_mem = nullptr; size_t win_bytes = 0;
  void *peer_base[VCL_MAX_PEERS] = {nullptr};
  PeerWindow hwin; PeerWindow *d_win = nullptr;
  HaloPush push; HaloPush *d_push = nullptr;     // d_push: device copy read by the product kernel's push phase
  bool fused_push = false;           // every destination's send list is one contiguous range: cg_update_kernel pushes
  long long push_lo[VCL_MAX_PUSH_RANGES] = {0}, push_hi[VCL_MAX_PUSH_RANGES] = {0};
  unsigned int wait_mask = 0;
  u32 *ord_start = nullptr, *ord_end = nullptr;   // row ranges of the blocks in the order [interior | boundary]
  u64 halo_seq = 0, red_seq = 0;     // exchanges EXECUTED so far (identical on every rank)
  int *d_err = nullptr;
  // ---- gather vector (peer-memory transport): [owned n | halo n_halo] doubles INSIDE the window; the CG keeps its search direction
  // p here and the neighbours push their boundary entries straight behind it, so the fused product gathers with ONE base ----
  double *gvec = nullptr;
  HaloPush push_g; HaloPush *d_push_g = nullptr;
  // ---- optional SELL-C copy of the slab (ViennaCLCUDADdist_csr_set_format): local column indices, sigma = 1 ----
  int fmt = 0, sell_C = 0, sell_passes = 0;
  u32 *s_cpb = nullptr, *s_bs = nullptr, *s_ci = nullptr; double *s_va = nullptr;
  unsigned char *s_needs = nullptr;  // per CTA pass of sell_kernel: does it gather from the halo?
};

namespace {

#define VCL_NCCL(b, api, expr) do { ncclResult_t r__ = (expr); if (r__ != ncclSuccess) return vcl_fail((b), ViennaCLB200CommError, (api)->GetErrorString(r__), __FILE__, __LINE__); } while (0)

__global__ void mark_external_kernel(u32 nnz, const u32 * __restrict__ ci, u32 rb, u32 re, unsigned int *bitmap)
{
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (size_t)gridDim.x * blockDim.x)
  {
    const u32 c = ci[k];
    if (c < rb || c >= re) atomicOr(&bitmap[c >> 5], 1u << (c & 31u));
  }
}

__global__ void remap_kernel(u32 nnz, const u32 * __restrict__ ci, u32 rb, u32 re, u32 n_local,
                             const u32 * __restrict__ halo_cols, int n_halo, u32 *out)
{
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (size_t)gridDim.x * blockDim.x)
  {
    const u32 c = ci[k];
    if (c >= rb && c < re) { out[k] = c - rb; continue; }
    int lo = 0, hi = n_halo - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (halo_cols[mid] < c) lo = mid + 1; else hi = mid; }
    out[k] = n_local + (u32)lo;
  }
}

__global__ void to_local_kernel(int cnt, u32 *idx, u32 rb)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) idx[i] -= rb;
}

// one CTA per row block: does any entry reference the halo?
__global__ void classify_blocks_kernel(int nblk, const u32 * __restrict__ blk, const u32 * __restrict__ rp,
                                       const u32 * __restrict__ ci_local, u32 n_local, int *flags)
{
  for (int b = blockIdx.x; b < nblk; b += gridDim.x)
  {
    const u32 k0 = rp[blk[b]], k1 = rp[blk[b + 1]];
    int found = 0;
    for (u32 k = k0 + threadIdx.x; k < k1; k += blockDim.x) found |= (ci_local[k] >= n_local);
    if (__syncthreads_or(found) && threadIdx.x == 0) flags[b] = 1;
  }
}

__global__ void pack_kernel(int cnt, const u32 * __restrict__ idx, const double * __restrict__ x, double *out)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) out[i] = x[idx[i]];
}

// x entries -> the neighbours' halo receive buffers (remote stores over NVLink); the CTA that finishes last publishes the
// sequence number to every destination with a release store.
__global__ void __launch_bounds__(256)
halo_push_kernel(HaloPush hp, const u32 * __restrict__ idx, const double * __restrict__ x, u64 seq, unsigned int *ticket,
                 const SolverState *st)
{
  __shared__ bool s_last;
  if (st != nullptr && st->done != VCL_RUNNING) return;
  const int par = (int)(seq & 1ULL);
  const int total = hp.begin[hp.ndst];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
  {
    int d = 0;
    while (i >= hp.begin[d + 1]) ++d;
    hp.dst[d][(size_t)par * hp.stride[d] + (size_t)(i - hp.begin[d])] = x[idx[i]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if ((int)threadIdx.x < hp.ndst) st_release_sys(hp.flag[threadIdx.x] + par * hp.W + hp.me, seq);
  if (threadIdx.x == 0) *ticket = 0u;
}

// stand-alone all-reduce of n <= 4 doubles (solver set-up)
__global__ void peer_allreduce_kernel(const PeerWindow *win, u64 seq, double *vals, int n)
{
  __shared__ double s_v[4], s_g[VCL_MAX_PEERS * 4];
  if (threadIdx.x < 4) s_v[threadIdx.x] = (int)threadIdx.x < n ? vals[threadIdx.x] : 0.0;
  peer_allreduce<4>(win, seq, s_v, s_g);
  __syncthreads();
  if ((int)threadIdx.x < n) vals[threadIdx.x] = s_v[threadIdx.x];
}

__global__ void cg_advance_kernel(SolverState *st)
{
  if (st->done == VCL_RUNNING) cg_advance(st);
}

// halo exchange of `x` (owned entries) into A->halo_buf, on the communication stream; compute stream is not blocked
ViennaCLStatus start_halo(ViennaCLBackend b, ViennaCLB200DistCsr A, const double *x)
{
  if (b->world == 1 || (A->n_halo == 0 && A->total_send == 0)) return ViennaCLSuccess;
  const NcclApi *api = vcl_nccl(nullptr);
  ncclComm_t comm = (ncclComm_t)b->nccl_comm;
  VCL_CUDA(b, cudaEventRecord(A->ev_x, b->stream));
  VCL_CUDA(b, cudaStreamWaitEvent(b->comm_stream, A->ev_x, 0));
  if (A->total_send > 0)
  {
    pack_kernel<<<std::min(vcl_div_up(A->total_send, 256), b->sm_count * 2), 256, 0, b->comm_stream>>>(A->total_send, A->send_idx, x, A->send_buf);
    VCL_LAUNCHED(b, "pack_kernel");
  }
  VCL_NCCL(b, api, api->GroupStart());
  for (int q = 0; q < b->world; ++q)
  {
    if (A->send_cnt[q] > 0) VCL_NCCL(b, api, api->Send(A->send_buf + A->send_off[q], (size_t)A->send_cnt[q], ncclDouble, q, comm, b->comm_stream));
    if (A->recv_cnt[q] > 0) VCL_NCCL(b, api, api->Recv(A->halo_buf + A->recv_off[q], (size_t)A->recv_cnt[q], ncclDouble, q, comm, b->comm_stream));
  }
  VCL_NCCL(b, api, api->GroupEnd());
  VCL_CUDA(b, cudaEventRecord(A->ev_halo, b->comm_stream));
  return ViennaCLSuccess;
}

ViennaCLStatus wait_halo(ViennaCLBackend b, ViennaCLB200DistCsr A)
{
  if (b->world == 1 || (A->n_halo == 0 && A->total_send == 0)) return ViennaCLSuccess;
  VCL_CUDA(b, cudaStreamWaitEvent(b->stream, A->ev_halo, 0));
  return ViennaCLSuccess;
}

// ---- peer-memory transport ----
ViennaCLStatus p2p_push(ViennaCLBackend b, ViennaCLB200DistCsr A, const double *x, u64 seq, const SolverState *st)
{
  if (A->push.ndst == 0) return ViennaCLSuccess;
  const int total = A->push.begin[A->push.ndst];
  const int grid = std::max(1, std::min(vcl_div_up(total, 256 * 4), b->sm_count));
  halo_push_kernel<<<grid, 256, 0, b->stream>>>(A->push, A->send_idx, x, seq, b->tickets + 8, st);
  VCL_LAUNCHED(b, "halo_push_kernel");
  return ViennaCLSuccess;
}

// all row blocks in one launch: interior first, boundary blocks (which wait for halo sequence number `seq`) last
// with_push: the kernel itself sends x's boundary entries before it starts on the row blocks (b->tickets + 8 counts the CTAs)
CsrDev p2p_all_blocks(ViennaCLBackend b, ViennaCLB200DistCsr A, u64 seq, bool with_push)
{
  const int par = (int)(seq & 1ULL);
  CsrDev d = {A->n, (u32)A->nnz, A->rp, A->ci_local, A->va, A->ord_start, A->ord_end, A->n_interior + A->n_boundary,
              A->n_interior, A->wait_mask, A->hwin.halo_flag[A->hwin.me] + par * A->hwin.W, seq, A->d_err,
              (with_push && A->push.ndst > 0) ? A->d_push : nullptr, A->send_idx, b->tickets + 8};
  return d;
}

// XS = false: the gathered vector and its halo are one contiguous array (the gather vector inside the window)
template<class Epi, bool XS = true>
ViennaCLStatus p2p_launch_csr(ViennaCLBackend b, const CsrDev &d, XVec xv, Epi epi)
{
  const int occ = vcl_occupancy(b, csr_stream_kernel<Epi, true, XS>, CSR_BLOCK_THREADS, CSR_SMEM_BYTES);
  const int grid = std::max(1, std::min(d.nblk, std::min(b->sm_count * occ, VCL_MAX_BLOCKS)));
  csr_stream_kernel<Epi, true, XS><<<grid, CSR_BLOCK_THREADS, CSR_SMEM_BYTES, b->stream>>>(d, xv, epi);
  VCL_LAUNCHED(b, "csr_stream_kernel(peer)");
  return ViennaCLSuccess;
}

XVec p2p_xvec(ViennaCLB200DistCsr A, const double *x, u64 seq)
{
  const PeerWindow &w = A->hwin;
  XVec xv = make_xvec(x, 0, 1, w.halo[w.me] + (size_t)(seq & 1ULL) * (size_t)w.halo_len[w.me], (u32)A->n);
  return xv;
}

// one CTA per pass of sell_kernel (256 / C slices): does any stored non-zero gather from the halo (local column >= n_local)?
__global__ void sell_pass_needs_halo_kernel(int passes, int spb, int nslices, int C, const u32 * __restrict__ cpb, const u32 * __restrict__ bs,
                                            const u32 * __restrict__ ci, const double * __restrict__ va, u32 n_local, unsigned char *needs)
{
  for (int p = blockIdx.x; p < passes; p += gridDim.x)
  {
    const int s0 = p * spb, s1 = min(s0 + spb, nslices);
    const u32 k0 = bs[s0], k1 = bs[s1 - 1] + cpb[s1 - 1] * (u32)C;
    int found = 0;
    for (u32 k = k0 + threadIdx.x; k < k1; k += blockDim.x) found |= (va[k] != 0.0 && ci[k] >= n_local);
    found = __syncthreads_or(found);
    if (threadIdx.x == 0) needs[p] = found ? 1 : 0;
  }
}

// the slab as a SELL matrix for sell_kernel<..., SPLIT = true>; in_kernel: halo pushes / flag waits inside the launch (peer-memory
// transport), otherwise the halo has arrived by stream order (NCCL transport, world 1) and the kernel neither pushes nor waits
SellDev dist_sell_dev(ViennaCLBackend b, ViennaCLB200DistCsr A, u64 seq, bool in_kernel, bool with_push)
{
  SellDev d = {A->n, A->sell_C, A->s_cpb, A->s_ci, A->s_bs, A->s_va, nullptr, A->s_needs, 0u, nullptr, seq, A->d_err, nullptr, nullptr, nullptr};
  if (in_kernel)
  {
    const int par = (int)(seq & 1ULL);
    d.wait_mask = A->wait_mask;
    d.wait_flags = A->hwin.halo_flag[A->hwin.me] + par * A->hwin.W;
    if (with_push && A->push.ndst > 0) { d.push = A->d_push; d.push_idx = A->send_idx; d.push_ticket = b->tickets + 8; }
  }
  return d;
}

// peer-memory transport: the row-partitioned product as ONE launch in the slab's format
template<class Epi>
ViennaCLStatus p2p_launch(ViennaCLBackend b, ViennaCLB200DistCsr A, u64 seq, const double *x, Epi epi, bool with_push)
{
  if (A->fmt == 1) return vcl_launch_sell_split(b, dist_sell_dev(b, A, seq, true, with_push), p2p_xvec(A, x, seq), epi);
  return p2p_launch_csr(b, p2p_all_blocks(b, A, seq, with_push), p2p_xvec(A, x, seq), epi);
}

ViennaCLStatus p2p_check(ViennaCLBackend b, ViennaCLB200DistCsr A)      // after a stream synchronisation
{
  int err = 0;
  VCL_CUDA(b, cudaMemcpy(&err, A->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (err)
  {
    char msg[192];
    std::snprintf(msg, sizeof msg, "peer-memory %s exchange timed out on rank %d: sequence number ...%u (of this object's chain), waiting for rank %u; "
                  "results from that launch on are undefined", ((unsigned)err >> 28) == 2u ? "reduction" : "halo", b->rank, (unsigned)err & 0xfffffu,
                  ((unsigned)err >> 20) & 0xffu);
    return vcl_fail(b, ViennaCLB200CommError, msg, __FILE__, __LINE__);
  }
  return ViennaCLSuccess;
}

CsrDev subset(ViennaCLB200DistCsr A, bool boundary)
{
  const int off = boundary ? A->n_interior : 0;
  CsrDev d = {A->n, (u32)A->nnz, A->rp, A->ci_local, A->va, A->ord_start + off, A->ord_end + off,
              boundary ? A->n_boundary : A->n_interior};
  d.wait_from = boundary ? 0 : d.nblk;                    // list positions >= wait_from may reference halo columns (nothing to wait for here:
  return d;                                               // the halo has arrived by stream order); interior blocks gather with one base
}

ViennaCLStatus allreduce_sum(ViennaCLBackend b, ViennaCLB200DistCsr A, double *buf, int count)
{
  if (b->world == 1 || count <= 0) return ViennaCLSuccess;
  if (A->p2p && count <= 4)                              // longer vectors (GMRES: up to krylov_dim dots at once) go through NCCL
  {
    peer_allreduce_kernel<<<1, 32, 0, b->stream>>>(A->d_win, ++A->red_seq, buf, count);
    VCL_LAUNCHED(b, "peer_allreduce_kernel");
    return ViennaCLSuccess;
  }
  const NcclApi *api = vcl_nccl(nullptr);
  VCL_NCCL(b, api, api->AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)b->nccl_comm, b->stream));
  return ViennaCLSuccess;
}

// y = A x with overlap; epilogue factory gives the interior / boundary epilogues
ViennaCLStatus dist_plain_prod(ViennaCLBackend b, ViennaCLB200DistCsr A, const double *x, double *y)
{
  EpiAxpby epi = {y, 0, 1, 1.0, 0.0};
  if (A->p2p)
  {
    // ONE launch: the product kernel pushes x's boundary entries from its own head, computes the interior row blocks while
    // the neighbours' entries travel, and waits for their flags only before the boundary blocks
    const u64 seq = ++A->halo_seq;
    if (getenv("VCL_B200_SEPARATE_PUSH"))                 // experiment knob: the round-1 form (push kernel + product kernel)
    {
      VCL_TRY(p2p_push(b, A, x, seq, nullptr));
      return p2p_launch_csr(b, p2p_all_blocks(b, A, seq, false), p2p_xvec(A, x, seq), epi);
    }
    return p2p_launch(b, A, seq, x, epi, true);
  }
  XVec xv = make_xvec(x, 0, 1, A->halo_buf, (u32)A->n);
  VCL_TRY(start_halo(b, A, x));
  if (A->fmt == 1)                                        // SELL slab on the NCCL transport / at world 1: no interior / boundary split
  {
    VCL_TRY(wait_halo(b, A));
    return vcl_launch_sell_split(b, dist_sell_dev(b, A, 0, false, false), xv, epi);
  }
  VCL_TRY(vcl_launch_csr_split(b, subset(A, false), xv, epi, b->stream));
  VCL_TRY(wait_halo(b, A));
  VCL_TRY(vcl_launch_csr_split(b, subset(A, true), xv, epi, b->stream));
  return ViennaCLSuccess;
}

__global__ void bicgstab_advance_kernel(SolverState *st)
{
  if (st->done == VCL_RUNNING) bicgstab_advance(st);
}

__global__ void bicgstab_maxit_kernel(SolverState *st)
{
  if (st->done == VCL_RUNNING && st->iters >= st->maxit) st->done = VCL_MAXIT;
}

__global__ void pcg_advance_kernel(SolverState *st, const double *delta)
{
  if (st->done == VCL_RUNNING) pcg_advance(st, *delta);
}

// NCCL transport (also the single-rank form): global sums of `count` rank-local totals, out of place so that re-issuing the
// collective after convergence (kernels skipped, `loc` unchanged) reproduces the same global sums
ViennaCLStatus nccl_sums(ViennaCLBackend b, const double *loc, double *glob, int count)
{
  if (b->world == 1)
  {
    VCL_CUDA(b, cudaMemcpyAsync(glob, loc, sizeof(double) * count, cudaMemcpyDeviceToDevice, b->stream));
    return ViennaCLSuccess;
  }
  const NcclApi *api = vcl_nccl(nullptr);
  VCL_NCCL(b, api, api->AllReduce(loc, glob, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)b->nccl_comm, b->stream));
  return ViennaCLSuccess;
}

// NCCL transport: y = A x [./ diag] as interior launch + halo wait + boundary launch; the rank-local totals <y,y>, <x,y>, <y,r0*>
// land in *o0, *o1, *o2 (any of them may be NULL)
template<bool USE_R0, bool JACOBI>
ViennaCLStatus nccl_fused_prod(ViennaCLBackend b, ViennaCLB200DistCsr A, const double *x, double *y, const double *r0, const double *diag,
                               SolverState *st, double *o0, double *o1, double *o2)
{
  XVec xv = make_xvec(x, 0, 1, A->halo_buf, (u32)A->n);
  VCL_TRY(start_halo(b, A, x));
  if (A->fmt == 1)
  {
    VCL_TRY(wait_halo(b, A));
    EpiFused<STEP_NONE, USE_R0, JACOBI> e = {y, x, r0, diag, b->partials, b->tickets, st, o0, o1, o2, {0.0, 0.0, 0.0}, nullptr};
    return vcl_launch_sell_split(b, dist_sell_dev(b, A, 0, false, false), xv, e);
  }
  const bool both = A->n_interior > 0 && A->n_boundary > 0;
  if (A->n_interior > 0)
  {
    EpiFused<STEP_NONE, USE_R0, JACOBI> ei = {y, x, r0, diag, b->partials, b->tickets, st, both ? A->tmp_sums + 0 : o0, both ? A->tmp_sums + 1 : o1,
                                              both ? A->tmp_sums + 2 : o2, {0.0, 0.0, 0.0}, nullptr};
    VCL_TRY(vcl_launch_csr_split(b, subset(A, false), xv, ei, b->stream));
  }
  VCL_TRY(wait_halo(b, A));
  if (A->n_boundary > 0)
  {
    EpiFused<STEP_NONE, USE_R0, JACOBI> eb = {y, x, r0, diag, b->partials, b->tickets, st, o0, o1, o2, {0.0, 0.0, 0.0}, both ? A->tmp_sums : nullptr};
    VCL_TRY(vcl_launch_csr_split(b, subset(A, true), xv, eb, b->stream));
  }
  return ViennaCLSuccess;
}

ViennaCLStatus dist_pull_state(ViennaCLBackend b, ViennaCLB200DistCsr A)
{
  VCL_CUDA(b, cudaMemcpyAsync(VCL_HSTATE(b), VCL_DSTATE(b), sizeof(SolverState), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  if (A->p2p) return p2p_check(b, A);
  return vcl_comm_check(b);                                // asynchronous NCCL errors (a peer died, a transport failed) surface here
}

// ------------------------------------------------------------------------------------------------
// Row-partitioned CG with a diagonal preconditioner (Jacobi / row scaling): the single-reduction PCG of solvers.cu (pcg_jacobi)
// over slabs -- per iteration one vector kernel, one halo exchange, one fused product whose last CTA all-reduces
// {gamma = <r,u>, delta = <w,u>} across the ranks and advances alpha / beta (peer-memory transport).
// ------------------------------------------------------------------------------------------------
ViennaCLStatus dist_pcg(ViennaCLBackend b, ViennaCLB200DistCsr A, const double *rhs, double *x, ViennaCLB200SolverTag *tag, int info_option)
{
  const long long n = A->n;
  const size_t need = ((size_t)std::max<long long>(n, 1) * sizeof(double) + 255) / 256 * 256;
  VCL_TRY(vcl_ws_reserve(b, 6 * need));
  char *w0 = (char*)b->ws;
  double *r = (double*)w0, *u = (double*)(w0 + need), *w = (double*)(w0 + 2 * need), *p = (double*)(w0 + 3 * need), *s = (double*)(w0 + 4 * need),
         *diag = (double*)(w0 + 5 * need);
  const int grid = (int)std::max(1LL, std::min((n / 2 + VEC_THREADS - 1) / VEC_THREADS, (long long)std::min(b->sm_count * 8, VCL_MAX_BLOCKS)));
  double *loc = b->dscal + 32;                             // [0] gamma, [1] delta (rank-local)

  // the diagonal entry of local row i is the entry with LOCAL column i (owned columns are renumbered to [0, n))
  if (n > 0) VCL_TRY(ViennaCLCUDADcsr_row_info(b, (int)n, A->rp, A->ci_local, A->va, diag, info_option));
  VCL_CUDA(b, cudaMemsetAsync(x, 0, sizeof(double) * n, b->stream));
  VCL_CUDA(b, cudaMemsetAsync(p, 0, sizeof(double) * n, b->stream));
  VCL_CUDA(b, cudaMemsetAsync(s, 0, sizeof(double) * n, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(r, rhs, sizeof(double) * n, cudaMemcpyDeviceToDevice, b->stream));
  VCL_CUDA(b, cudaMemsetAsync(b->dscal, 0, 2 * sizeof(double), b->stream));
  VCL_CUDA(b, cudaMemsetAsync(loc, 0, 2 * sizeof(double), b->stream));
  if (n > 0)
  {
    pcg_init_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, r, u, diag, b->partials, b->tickets, b->dscal + 0);
    VCL_LAUNCHED(b, "pcg_init_kernel");
  }
  VCL_TRY(dist_plain_prod(b, A, u, w));
  if (n > 0) VCL_TRY(vcl_dot_async(b, n, w, 0, 1, u, 0, 1, b->dscal + 1));
  VCL_TRY(allreduce_sum(b, A, b->dscal, 2));
  VCL_CUDA(b, cudaMemcpyAsync(b->hscal, b->dscal, 2 * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  const double gamma0 = b->hscal[0], delta0 = b->hscal[1];
  if (std::fabs(gamma0) <= tag->abs_tolerance * tag->abs_tolerance) return ViennaCLSuccess;

  SolverState *h = VCL_HSTATE(b);
  std::memset(h, 0, sizeof(SolverState));
  h->alpha = gamma0 / delta0; h->beta = 0.0; h->ip_rr0 = gamma0; h->norm_rhs_sq = gamma0; h->norm_rhs = std::sqrt(std::fabs(gamma0));
  h->tol = tag->tolerance; h->abs_tol = tag->abs_tolerance; h->maxit = tag->max_iterations; h->sums[0] = gamma0;
  VCL_CUDA(b, cudaMemcpyAsync(b->dstate, h, sizeof(SolverState), cudaMemcpyHostToDevice, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  SolverState *st = VCL_DSTATE(b);

  const int kBatch = 32;
  int launched = 0;
  const u64 halo_base = A->halo_seq, red_base = A->red_seq;
  while (launched < tag->max_iterations)
  {
    const int nb = std::min(kBatch, tag->max_iterations - launched);
    VCL_RANGE("vcl:batch");
    for (int k = 0; k < nb; ++k)
    {
      if (n > 0)
      {
        pcg_update_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, x, r, u, w, p, s, diag, st, b->partials, b->tickets, loc + 0);
        VCL_LAUNCHED(b, "pcg_update_kernel");
      }
      if (A->p2p)
      {
        const u64 hseq = halo_base + (u64)(launched + k + 1), rseq = red_base + (u64)(launched + k + 1);
        EpiFused<STEP_NONE, false, false> e = {w, u, nullptr, nullptr, b->partials, b->tickets, st, nullptr, nullptr, nullptr,
                                               {0.0, 0.0, 0.0}, nullptr, A->d_win, rseq, loc + 0, DIST_PCG};
        VCL_TRY(p2p_launch(b, A, hseq, u, e, true));
      }
      else
      {
        VCL_TRY((nccl_fused_prod<false, false>(b, A, u, w, nullptr, nullptr, st, nullptr, loc + 1, nullptr)));
        VCL_TRY(nccl_sums(b, loc, b->dscal + 36, 2));
        VCL_CUDA(b, cudaMemcpyAsync(&st->sums[0], b->dscal + 36, sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
        pcg_advance_kernel<<<1, 1, 0, b->stream>>>(st, b->dscal + 37);
        VCL_LAUNCHED(b, "pcg_advance_kernel");
      }
    }
    launched += nb;
    VCL_TRY(dist_pull_state(b, A));
    if (h->done != VCL_RUNNING) break;
  }
  if (A->p2p) { A->halo_seq = halo_base + (u64)h->iters; A->red_seq = red_base + (u64)h->iters; }
  tag->iters = h->iters;
  tag->error = std::sqrt(std::fabs(h->sums[0] / gamma0));
  return ViennaCLSuccess;
}

// ---- peer-memory transport set-up: allocate the window, exchange IPC handles (NCCL all-gather), map the peers ----
// M[p*W + q] = number of halo entries rank p receives from rank q (known to every rank).
static size_t win_off_rflag(int W) { return (size_t)2 * W * sizeof(u64); }
static size_t win_off_red(int W)   { return (size_t)4 * W * sizeof(u64); }
static size_t win_off_halo(int W)  { return ((size_t)4 * W * sizeof(u64) + (size_t)8 * W * sizeof(double) + 255) / 256 * 256; }

struct PeerHello { cudaIpcMemHandle_t handle; long long n_halo; long long n; int device; int ok; };
static size_t win_off_gvec(int W, long long n_halo) { return (win_off_halo(W) + (size_t)2 * (size_t)std::max<long long>(n_halo, 1) * sizeof(double) + 255) / 256 * 256; }

ViennaCLStatus setup_p2p(ViennaCLBackend b, ViennaCLB200DistCsr A, const std::vector<int> &M)
{
  const int W = b->world, me = b->rank;
  const NcclApi *api = vcl_nccl(nullptr);
  ncclComm_t comm = (ncclComm_t)b->nccl_comm;
  const char *env = getenv("VCL_B200_DIST_TRANSPORT");
  int want = (W <= VCL_MAX_PEERS) && !(env && std::string(env) == "nccl");

  // own window (>= 2 MiB so that the allocation is not carved out of a shared driver block)
  A->win_bytes = std::max<size_t>(win_off_gvec(W, A->n_halo) + ((size_t)A->n + (size_t)std::max(A->n_halo, 1)) * sizeof(double), (size_t)2 << 20);
  PeerHello hello;
  std::memset(&hello, 0, sizeof(hello));
  hello.n_halo = A->n_halo; hello.n = A->n; hello.device = b->device; hello.ok = want;
  if (want)
  {
    if (cudaMalloc(&A->win_mem, A->win_bytes) != cudaSuccess || cudaMemsetAsync(A->win_mem, 0, A->win_bytes, b->stream) != cudaSuccess ||
        cudaIpcGetMemHandle(&hello.handle, A->win_mem) != cudaSuccess)
    { cudaGetLastError(); hello.ok = 0; }
  }
  // all-gather the hellos (byte-wise through NCCL; set-up only)
  PeerHello *d_h = nullptr;
  std::vector<PeerHello> all(W);
  VCL_CUDA(b, cudaMalloc(&d_h, sizeof(PeerHello) * (W + 1)));
  VCL_CUDA(b, cudaMemcpyAsync(d_h + W, &hello, sizeof(PeerHello), cudaMemcpyHostToDevice, b->stream));
  VCL_NCCL(b, api, api->AllGather(d_h + W, d_h, sizeof(PeerHello), ncclChar, comm, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(all.data(), d_h, sizeof(PeerHello) * W, cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  int ok = 1;
  for (int q = 0; q < W; ++q) ok &= all[q].ok;
  if (ok)
  {
    for (int q = 0; q < W && ok; ++q)
    {
      if (q == me) { A->peer_base[q] = A->win_mem; continue; }
      if (cudaIpcOpenMemHandle(&A->peer_base[q], all[q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
      { cudaGetLastError(); A->peer_base[q] = nullptr; ok = 0; }
    }
  }
  // everybody must agree (and this collective is also the barrier between "windows zeroed" and the first push)
  int *d_ok = reinterpret_cast<int*>(d_h);
  VCL_CUDA(b, cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, b->stream));
  VCL_NCCL(b, api, api->AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, comm, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  VCL_CUDA(b, cudaFree(d_h));
  if (!ok)
  {
    for (int q = 0; q < W; ++q) if (q != me && A->peer_base[q]) { cudaIpcCloseMemHandle(A->peer_base[q]); A->peer_base[q] = nullptr; }
    if (A->win_mem) { cudaFree(A->win_mem); A->win_mem = nullptr; }
    if (want && me == 0) fprintf(stderr, "libvcl_b200: peer-memory (CUDA IPC) mapping unavailable, using the NCCL transport\n");
    A->p2p = false;
    return ViennaCLSuccess;
  }

  PeerWindow &w = A->hwin;
  std::memset(&w, 0, sizeof(w));
  w.W = W; w.me = me;
  for (int q = 0; q < W; ++q)
  {
    char *base = static_cast<char*>(A->peer_base[q]);
    w.halo_flag[q] = reinterpret_cast<u64*>(base);
    w.red_flag[q]  = reinterpret_cast<u64*>(base + win_off_rflag(W));
    w.red[q]       = reinterpret_cast<double*>(base + win_off_red(W));
    w.halo[q]      = reinterpret_cast<double*>(base + win_off_halo(W));
    w.halo_len[q]  = std::max<long long>(all[q].n_halo, 1);
  }
  VCL_CUDA(b, cudaMalloc(&A->d_err, sizeof(int)));
  VCL_CUDA(b, cudaMemset(A->d_err, 0, sizeof(int)));
  w.err = A->d_err;
#ifdef VCL_PEER_DEBUG
  VCL_CUDA(b, cudaMalloc(&w.dbg, sizeof(u64) * 4096));
  VCL_CUDA(b, cudaMemset(w.dbg, 0, sizeof(u64) * 4096));
#endif
  VCL_CUDA(b, cudaMalloc(&A->d_win, sizeof(PeerWindow)));
  VCL_CUDA(b, cudaMemcpy(A->d_win, &w, sizeof(PeerWindow), cudaMemcpyHostToDevice));

  // destinations: every rank with traffic in EITHER direction (symmetric pairs bound how far a rank can run ahead)
  HaloPush &hp = A->push;
  std::memset(&hp, 0, sizeof(hp));
  hp.me = me; hp.W = W;
  A->wait_mask = 0;
  for (int q = 0; q < W; ++q)
  {
    if (q == me || (A->send_cnt[q] == 0 && A->recv_cnt[q] == 0)) continue;
    const int d = hp.ndst++;
    hp.begin[d] = A->send_off[q];
    long long off_in_q = 0;                      // where my segment starts in q's halo: after the segments of ranks < me
    for (int p2 = 0; p2 < me; ++p2) off_in_q += M[(size_t)q * W + p2];
    hp.dst[d] = w.halo[q] + off_in_q;
    hp.stride[d] = w.halo_len[q];
    hp.flag[d] = w.halo_flag[q];
    A->wait_mask |= 1u << q;
  }
  // send segments are contiguous and ordered by destination rank; ranks without entries contribute empty segments
  hp.begin[hp.ndst] = A->total_send;
  VCL_CUDA(b, cudaMalloc(&A->d_push, sizeof(HaloPush)));
  VCL_CUDA(b, cudaMemcpy(A->d_push, &hp, sizeof(HaloPush), cudaMemcpyHostToDevice));
  // the same destinations for the gather-vector form: rank q's halo part starts n_q entries into its gather vector; ONE buffer
  // (parity stride 0) -- safe inside the CG, where an all-reduce separates consecutive exchanges (see ViennaCLCUDADdist_csr_cg)
  A->gvec = reinterpret_cast<double*>(static_cast<char*>(A->win_mem) + win_off_gvec(W, A->n_halo));
  A->push_g = hp;
  {
    int d = 0;
    for (int q = 0; q < W; ++q)
    {
      if (q == me || (A->send_cnt[q] == 0 && A->recv_cnt[q] == 0)) continue;
      long long off_in_q = 0;
      for (int p2 = 0; p2 < me; ++p2) off_in_q += M[(size_t)q * W + p2];
      double *gq = reinterpret_cast<double*>(static_cast<char*>(A->peer_base[q]) + win_off_gvec(W, all[q].n_halo));
      A->push_g.dst[d] = gq + all[q].n + off_in_q;
      A->push_g.stride[d] = 0;
      ++d;
    }
  }
  VCL_CUDA(b, cudaMalloc(&A->d_push_g, sizeof(HaloPush)));
  VCL_CUDA(b, cudaMemcpy(A->d_push_g, &A->push_g, sizeof(HaloPush), cudaMemcpyHostToDevice));
  // contiguous send ranges?  (slab partitions of banded matrices: yes) -> the producing kernel can push by itself
  A->fused_push = false;
  if (hp.ndst > 0 && hp.ndst <= VCL_MAX_PUSH_RANGES && !(getenv("VCL_B200_NO_FUSED_PUSH")))
  {
    std::vector<u32> idx((size_t)std::max(A->total_send, 1));
    if (A->total_send) VCL_CUDA(b, cudaMemcpy(idx.data(), A->send_idx, sizeof(u32) * A->total_send, cudaMemcpyDeviceToHost));
    bool contiguous = true;
    for (int d = 0; d < hp.ndst && contiguous; ++d)
    {
      const int b0 = hp.begin[d], b1 = hp.begin[d + 1];
      A->push_lo[d] = b1 > b0 ? (long long)idx[b0] : 0;
      A->push_hi[d] = A->push_lo[d] + (b1 - b0);
      for (int i = b0; i < b1; ++i) if (idx[i] != idx[b0] + (u32)(i - b0)) { contiguous = false; break; }
    }
    // Pushing from cg_update_kernel (the kernel that PRODUCES p) was the round-1 default; pushing from the head of the consuming
    // product kernel measured faster (512^3 CG: 2169 -> 2197 it/s on 8 GPUs, 589.9 -> 592.4 on 2; profiles/ab_headpush_r2u.log): the
    // update kernel runs at its single-GPU speed and the pushed entries still arrive long before the boundary blocks need them.
    // VCL_B200_FUSED_PUSH=1 selects the old form.
    A->fused_push = contiguous && getenv("VCL_B200_FUSED_PUSH") != nullptr;
  }
  A->p2p = true;
  return ViennaCLSuccess;
}

} // namespace

extern "C" {

ViennaCLStatus ViennaCLCUDADdist_csr_create(ViennaCLBackend b, long long global_rows, long long row_begin, long long row_end,
                                            ViennaCLInt local_nnz, const unsigned int *row_ptr, const unsigned int *col_idx_global,
                                            const double *values, ViennaCLB200DistCsr *out)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, out && global_rows > 0 && row_begin >= 0 && row_begin <= row_end && row_end <= global_rows, "bad row range");
  VCL_REQUIRE(b, global_rows <= 0xFFFFFFFFLL && row_end - row_begin <= 0x7FFFFFFFLL, "32-bit indices (compressed_matrix.hpp:1190-1197)");
  VCL_REQUIRE(b, row_ptr && (local_nnz == 0 || (col_idx_global && values)), "null pointer");
  VCL_REQUIRE(b, b->world == 1 || b->nccl_comm, "call ViennaCLBackendCommInit first");
  VCL_CUDA(b, cudaSetDevice(b->device));
  const NcclApi *api = b->world > 1 ? vcl_nccl(nullptr) : nullptr;
  ncclComm_t comm = (ncclComm_t)b->nccl_comm;
  const int W = b->world, me = b->rank;

  ViennaCLB200DistCsr A = new ViennaCLB200DistCsr_impl();
  A->global_rows = global_rows; A->rb = row_begin; A->re = row_end; A->n = (int)(row_end - row_begin); A->nnz = local_nnz;
  A->rp = row_ptr; A->ci_global = col_idx_global; A->va = values;
  A->recv_cnt.assign(W, 0); A->recv_off.assign(W, 0); A->send_cnt.assign(W, 0); A->send_off.assign(W, 0);
  VCL_CUDA(b, cudaEventCreateWithFlags(&A->ev_x, cudaEventDisableTiming));
  VCL_CUDA(b, cudaEventCreateWithFlags(&A->ev_halo, cudaEventDisableTiming));
  VCL_CUDA(b, cudaMalloc(&A->tmp_sums, 4 * sizeof(double)));

  // ---- 1. row ranges of all ranks ----
  std::vector<long long> starts(W + 1, 0);
  if (W > 1)
  {
    long long *d_all = nullptr;
    VCL_CUDA(b, cudaMalloc(&d_all, sizeof(long long) * (W + 1)));
    VCL_CUDA(b, cudaMemcpyAsync(d_all + W, &row_begin, sizeof(long long), cudaMemcpyHostToDevice, b->stream));
    VCL_NCCL(b, api, api->AllGather(d_all + W, d_all, 1, ncclInt64, comm, b->stream));
    VCL_CUDA(b, cudaMemcpyAsync(starts.data(), d_all, sizeof(long long) * W, cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    VCL_CUDA(b, cudaFree(d_all));
  }
  starts[W] = global_rows;
  if (W == 1) starts[0] = 0;
  VCL_REQUIRE(b, starts[me] == row_begin && starts[me + 1] == row_end, "row ranges of the ranks must be contiguous and ordered by rank");

  // ---- 2. halo columns: bitmap on the device, compaction on the host (set-up path) ----
  std::vector<u32> halo;
  if (W > 1 && local_nnz > 0)
  {
    const size_t words = (size_t)(global_rows + 31) / 32;
    unsigned int *bitmap = nullptr;
    VCL_CUDA(b, cudaMalloc(&bitmap, words * sizeof(unsigned int)));
    VCL_CUDA(b, cudaMemsetAsync(bitmap, 0, words * sizeof(unsigned int), b->stream));
    mark_external_kernel<<<b->sm_count * 8, 256, 0, b->stream>>>((u32)local_nnz, col_idx_global, (u32)row_begin, (u32)row_end, bitmap);
    VCL_LAUNCHED(b, "mark_external_kernel");
    std::vector<unsigned int> hb(words);
    VCL_CUDA(b, cudaMemcpyAsync(hb.data(), bitmap, words * sizeof(unsigned int), cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    VCL_CUDA(b, cudaFree(bitmap));
    for (size_t w = 0; w < words; ++w)
    {
      unsigned int bits = hb[w];
      while (bits) { const int t = __builtin_ctz(bits); halo.push_back((u32)(w * 32 + t)); bits &= bits - 1; }
    }
  }
  A->n_halo = (int)halo.size();
  {
    int q = 0, off = 0;
    for (size_t i = 0; i < halo.size(); ++i)
    {
      while (halo[i] >= (u32)starts[q + 1]) ++q;
      A->recv_cnt[q]++;
    }
    for (int p = 0; p < W; ++p) { A->recv_off[p] = off; off += A->recv_cnt[p]; }
  }

  // ---- 3. renumber columns to [owned | halo] ----
  u32 *d_halo = nullptr;
  if (W > 1)
  {
    VCL_CUDA(b, cudaMalloc(&A->ci_local, sizeof(u32) * std::max(local_nnz, 1)));
    VCL_CUDA(b, cudaMalloc(&d_halo, sizeof(u32) * std::max(A->n_halo, 1)));
    VCL_CUDA(b, cudaMalloc(&A->halo_buf, sizeof(double) * std::max(A->n_halo, 1)));
    if (A->n_halo) VCL_CUDA(b, cudaMemcpyAsync(d_halo, halo.data(), sizeof(u32) * A->n_halo, cudaMemcpyHostToDevice, b->stream));
    if (local_nnz > 0)
    {
      remap_kernel<<<b->sm_count * 8, 256, 0, b->stream>>>((u32)local_nnz, col_idx_global, (u32)row_begin, (u32)row_end, (u32)A->n,
                                                          d_halo, A->n_halo, A->ci_local);
      VCL_LAUNCHED(b, "remap_kernel");
    }
  }
  else
    A->ci_local = const_cast<u32*>(col_idx_global);      // single rank: global == local numbering

  // ---- 4. send lists: everyone learns what every peer needs from it ----
  std::vector<int> M((size_t)W * W, 0);
  if (W > 1)
  {
    int *d_cnt = nullptr;
    VCL_CUDA(b, cudaMalloc(&d_cnt, sizeof(int) * (size_t)W * (W + 1)));
    VCL_CUDA(b, cudaMemcpyAsync(d_cnt + (size_t)W * W, A->recv_cnt.data(), sizeof(int) * W, cudaMemcpyHostToDevice, b->stream));
    VCL_NCCL(b, api, api->AllGather(d_cnt + (size_t)W * W, d_cnt, (size_t)W, ncclInt32, comm, b->stream));
    VCL_CUDA(b, cudaMemcpyAsync(M.data(), d_cnt, sizeof(int) * (size_t)W * W, cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    VCL_CUDA(b, cudaFree(d_cnt));
    int off = 0;
    for (int p = 0; p < W; ++p) { A->send_cnt[p] = M[(size_t)p * W + me]; A->send_off[p] = off; off += A->send_cnt[p]; }
    A->total_send = off;
    VCL_CUDA(b, cudaMalloc(&A->send_idx, sizeof(u32) * std::max(off, 1)));
    VCL_CUDA(b, cudaMalloc(&A->send_buf, sizeof(double) * std::max(off, 1)));
    VCL_NCCL(b, api, api->GroupStart());
    for (int q = 0; q < W; ++q)
    {
      if (A->recv_cnt[q] > 0) VCL_NCCL(b, api, api->Send(d_halo + A->recv_off[q], (size_t)A->recv_cnt[q], ncclUint32, q, comm, b->stream));
      if (A->send_cnt[q] > 0) VCL_NCCL(b, api, api->Recv(A->send_idx + A->send_off[q], (size_t)A->send_cnt[q], ncclUint32, q, comm, b->stream));
    }
    VCL_NCCL(b, api, api->GroupEnd());
    if (off > 0)
    {
      to_local_kernel<<<std::min(vcl_div_up(off, 256), b->sm_count * 4), 256, 0, b->stream>>>(off, A->send_idx, (u32)row_begin);
      VCL_LAUNCHED(b, "to_local_kernel");
    }
  }

  // ---- 5. row blocks and their interior / boundary split ----
  if (A->n > 0)
  {
    // the plan of the single-domain kernels (device-side for evenly filled rows), then its block list on the host
    ViennaCLInt nb = 0;
    VCL_TRY(ViennaCLCUDAcsr_row_blocks(b, A->n, row_ptr, nullptr, &nb));
    A->nblk = nb;
    VCL_CUDA(b, cudaMalloc(&A->blk, sizeof(u32) * ((size_t)nb + 1)));
    VCL_TRY(ViennaCLCUDAcsr_row_blocks(b, A->n, row_ptr, A->blk, &nb));
    std::vector<u32> blk((size_t)nb + 1);
    VCL_CUDA(b, cudaMemcpyAsync(blk.data(), A->blk, sizeof(u32) * blk.size(), cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    std::vector<int> flags(A->nblk, 0);
    if (W > 1 && A->n_halo > 0)
    {
      int *d_flags = nullptr;
      VCL_CUDA(b, cudaMalloc(&d_flags, sizeof(int) * A->nblk));
      VCL_CUDA(b, cudaMemsetAsync(d_flags, 0, sizeof(int) * A->nblk, b->stream));
      classify_blocks_kernel<<<std::min(A->nblk, b->sm_count * 8), 256, 0, b->stream>>>(A->nblk, A->blk, row_ptr, A->ci_local, (u32)A->n, d_flags);
      VCL_LAUNCHED(b, "classify_blocks_kernel");
      VCL_CUDA(b, cudaMemcpyAsync(flags.data(), d_flags, sizeof(int) * A->nblk, cudaMemcpyDeviceToHost, b->stream));
      VCL_CUDA(b, cudaStreamSynchronize(b->stream));
      VCL_CUDA(b, cudaFree(d_flags));
    }
    std::vector<u32> in, bd;
    for (int i = 0; i < A->nblk; ++i) (flags[i] ? bd : in).push_back((u32)i);
    A->n_interior = (int)in.size(); A->n_boundary = (int)bd.size();
    // row ranges in the order [interior | boundary]: the kernels read them without an indirection
    std::vector<u32> os, oe;
    for (u32 i : in) { os.push_back(blk[i]); oe.push_back(blk[i + 1]); }
    for (u32 i : bd) { os.push_back(blk[i]); oe.push_back(blk[i + 1]); }
    VCL_CUDA(b, cudaMalloc(&A->ord_start, sizeof(u32) * std::max<size_t>(os.size(), 1)));
    VCL_CUDA(b, cudaMalloc(&A->ord_end, sizeof(u32) * std::max<size_t>(oe.size(), 1)));
    VCL_CUDA(b, cudaMemcpyAsync(A->ord_start, os.data(), sizeof(u32) * os.size(), cudaMemcpyHostToDevice, b->stream));
    VCL_CUDA(b, cudaMemcpyAsync(A->ord_end, oe.data(), sizeof(u32) * oe.size(), cudaMemcpyHostToDevice, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  }
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  if (d_halo) VCL_CUDA(b, cudaFree(d_halo));
  if (W > 1) VCL_TRY(setup_p2p(b, A, M));
  *out = A;
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDADdist_csr_destroy(ViennaCLBackend b, ViennaCLB200DistCsr *pA)
{
  VCL_CHECK_BACKEND(b);
  if (!pA || !*pA) return ViennaCLSuccess;
  ViennaCLB200DistCsr A = *pA;
  cudaStreamSynchronize(b->stream); cudaStreamSynchronize(b->comm_stream);
  if (A->ci_local && A->ci_local != A->ci_global) cudaFree(A->ci_local);
  cudaFree(A->send_idx); cudaFree(A->send_buf); cudaFree(A->halo_buf); cudaFree(A->blk); cudaFree(A->ord_start); cudaFree(A->ord_end);
  cudaFree(A->tmp_sums);
  cudaFree(A->s_cpb); cudaFree(A->s_bs); cudaFree(A->s_ci); cudaFree(A->s_va); cudaFree(A->s_needs);
  if (A->p2p)
  {
    // nobody may unmap or free a window that a partner is still writing to
    const NcclApi *api = vcl_nccl(nullptr);
    if (api && b->nccl_comm)
    {
      api->AllReduce(A->d_err, A->d_err, 1, ncclInt32, ncclMax, (ncclComm_t)b->nccl_comm, b->stream);
      cudaStreamSynchronize(b->stream);
    }
    for (int q = 0; q < b->world; ++q) if (q != b->rank && A->peer_base[q]) cudaIpcCloseMemHandle(A->peer_base[q]);
    cudaFree(A->win_mem); cudaFree(A->d_win); cudaFree(A->d_err); cudaFree(A->d_push); cudaFree(A->d_push_g);
  }
  if (A->ev_x) cudaEventDestroy(A->ev_x);
  if (A->ev_halo) cudaEventDestroy(A->ev_halo);
  delete A;
  *pA = nullptr;
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDADdist_csr_info(ViennaCLBackend b, ViennaCLB200DistCsr A, ViennaCLInt *peer_memory, ViennaCLInt *halo_entries,
                                          ViennaCLInt *interior_blocks, ViennaCLInt *boundary_blocks)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, A, "null matrix");
  if (peer_memory) *peer_memory = A->p2p ? 1 : 0;
  if (halo_entries) *halo_entries = A->n_halo;
  if (interior_blocks) *interior_blocks = A->n_interior;
  if (boundary_blocks) *boundary_blocks = A->n_boundary;
  return ViennaCLSuccess;
}

// Storage format of the slab for the products and solver steps: 0 = CSR (default), 1 = SELL-C (sigma = 1, the reference's layout,
// sliced_ell_matrix.hpp:134-214) built on the device from the slab with LOCAL column indices.  Per-row arithmetic is then the
// reference's SELL arithmetic (fused multiply-adds), i.e. results equal the single-domain SELL product bit for bit.
ViennaCLStatus ViennaCLCUDADdist_csr_set_format(ViennaCLBackend b, ViennaCLB200DistCsr A, ViennaCLInt format, ViennaCLInt rows_per_block)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, A && (format == 0 || format == 1), "format: 0 (CSR) or 1 (SELL)");
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  cudaFree(A->s_cpb); cudaFree(A->s_bs); cudaFree(A->s_ci); cudaFree(A->s_va); cudaFree(A->s_needs);
  A->s_cpb = A->s_bs = A->s_ci = nullptr; A->s_va = nullptr; A->s_needs = nullptr; A->fmt = 0;
  if (format == 0 || A->n == 0) return ViennaCLSuccess;
  const int C = rows_per_block > 0 ? rows_per_block : 32;
  VCL_REQUIRE(b, C <= 4096, "rows_per_block too large");
  const int nslices = (A->n - 1) / C + 1;
  VCL_CUDA(b, cudaMalloc(&A->s_cpb, sizeof(u32) * nslices));
  VCL_CUDA(b, cudaMalloc(&A->s_bs, sizeof(u32) * nslices));
  long long padded = 0;
  VCL_TRY(ViennaCLCUDADcsr2sell(b, A->n, C, A->rp, A->ci_local, A->va, A->s_cpb, A->s_bs, &padded, nullptr, nullptr));
  VCL_CUDA(b, cudaMalloc(&A->s_ci, sizeof(u32) * std::max<long long>(padded, 1)));
  VCL_CUDA(b, cudaMalloc(&A->s_va, sizeof(double) * std::max<long long>(padded, 1)));
  VCL_TRY(ViennaCLCUDADcsr2sell(b, A->n, C, A->rp, A->ci_local, A->va, A->s_cpb, A->s_bs, &padded, A->s_ci, A->s_va));
  const int spb = C <= CSR_BLOCK_THREADS ? CSR_BLOCK_THREADS / C : 1;
  A->sell_passes = (nslices + spb - 1) / spb;
  VCL_CUDA(b, cudaMalloc(&A->s_needs, (size_t)A->sell_passes));
  sell_pass_needs_halo_kernel<<<std::min(A->sell_passes, b->sm_count * 8), 256, 0, b->stream>>>(A->sell_passes, spb, nslices, C, A->s_cpb, A->s_bs, A->s_ci, A->s_va,
                                                                                                 (u32)A->n, A->s_needs);
  VCL_LAUNCHED(b, "sell_pass_needs_halo_kernel");
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  A->sell_C = C; A->fmt = 1;
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDADdist_csrmv(ViennaCLBackend b, ViennaCLB200DistCsr A, const double *x_local, double *y_local)
{
  VCL_RANGE("vcl:dist_csrmv");
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, A && (A->n == 0 || (x_local && y_local)) && x_local != y_local, "bad arguments");
  if (A->n == 0 && b->world == 1) return ViennaCLSuccess;
  return dist_plain_prod(b, A, x_local, y_local);
}

// cg.hpp:128-187 over row-partitioned data: local fused kernels + halo exchange + one allreduce per iteration.
ViennaCLStatus ViennaCLCUDADdist_csr_cg(ViennaCLBackend b, ViennaCLB200DistCsr A, const double *rhs, double *x, ViennaCLB200SolverTag *tag)
{
  VCL_RANGE("vcl:dist_cg");
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, A && tag, "bad arguments");
  VCL_REQUIRE(b, tag->monitor == nullptr, "monitor callbacks are not supported on the row-partitioned path");
  const long long n = A->n;
  tag->iters = 0; tag->error = 0.0;
  VCL_REQUIRE(b, n == 0 || (rhs && x), "null vector");
  VCL_CUDA(b, cudaSetDevice(b->device));
  {
    // diagonal preconditioners (jacobi_precond.hpp:103-130, row_scaling.hpp): row_info option 3 / 0 / 1 / 2
    int opt = -1;
    switch (tag->precond)
    {
    case ViennaCLB200PrecondJacobi: opt = 3; break;
    case ViennaCLB200PrecondRowScalingInf: opt = 0; break;
    case ViennaCLB200PrecondRowScaling1: opt = 1; break;
    case ViennaCLB200PrecondRowScaling2: opt = 2; break;
    default: break;
    }
    if (opt >= 0) return dist_pcg(b, A, rhs, x, tag, opt);
  }
  VCL_REQUIRE(b, tag->precond == ViennaCLB200PrecondNone, "CG: unknown preconditioner id");
  const size_t need = ((size_t)std::max<long long>(n, 1) * sizeof(double) + 255) / 256 * 256;
  VCL_TRY(vcl_ws_reserve(b, 3 * need));
  double *r = (double*)b->ws, *p = (double*)((char*)b->ws + need), *Ap = (double*)((char*)b->ws + 2 * need);
  // Gather-vector form (peer-memory transport, CSR slabs): p lives inside the window with its halo right behind it -- the neighbours
  // push their boundary entries of p straight to p[n ...] -- so the fused product gathers with the plain one-base addressing (the
  // [owned | halo] base selection cost ~12 us of a 285 us product).  ONE halo buffer instead of two: inside the CG consecutive
  // exchanges are separated by the all-reduce of the iteration between them -- a neighbour pushes exchange i+1 from the head of its
  // product i+1, i.e. after it has the all-reduced sums of iteration i, which contain this rank's contribution, which this rank's
  // last CTA sends only after ALL its CTAs have finished product i (and with it their reads of halo i).
  const bool use_g = A->p2p && A->fmt == 0 && !A->fused_push && A->gvec != nullptr && !getenv("VCL_B200_NO_GVEC");
  if (use_g) p = A->gvec;

  VCL_CUDA(b, cudaMemsetAsync(x, 0, sizeof(double) * n, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(r, rhs, sizeof(double) * n, cudaMemcpyDeviceToDevice, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(p, rhs, sizeof(double) * n, cudaMemcpyDeviceToDevice, b->stream));
  VCL_TRY(dist_plain_prod(b, A, p, Ap));
  VCL_CUDA(b, cudaMemsetAsync(b->dscal, 0, 3 * sizeof(double), b->stream));
  if (n > 0)
  {
    VCL_TRY(vcl_dot_async(b, n, r, 0, 1, r, 0, 1, b->dscal + 0));
    VCL_TRY(vcl_dot_async(b, n, p, 0, 1, Ap, 0, 1, b->dscal + 1));
    VCL_TRY(vcl_dot_async(b, n, Ap, 0, 1, Ap, 0, 1, b->dscal + 2));
  }
  VCL_TRY(allreduce_sum(b, A, b->dscal, 3));
  VCL_CUDA(b, cudaMemcpyAsync(b->hscal, b->dscal, 3 * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));

  double norm_rhs_squared = std::sqrt(b->hscal[0]); norm_rhs_squared *= norm_rhs_squared;
  if (norm_rhs_squared <= tag->abs_tolerance * tag->abs_tolerance) return ViennaCLSuccess;
  const double rr = norm_rhs_squared;
  const double alpha = rr / b->hscal[1];
  double beta = std::sqrt(b->hscal[2]); beta = (alpha * alpha * beta * beta - rr) / rr;

  SolverState *h = VCL_HSTATE(b);
  std::memset(h, 0, sizeof(SolverState));
  h->alpha = alpha; h->beta = beta; h->norm_rhs_sq = norm_rhs_squared; h->norm_rhs = std::sqrt(norm_rhs_squared);
  h->tol = tag->tolerance; h->abs_tol = tag->abs_tolerance; h->maxit = tag->max_iterations; h->sums[0] = rr;
  VCL_CUDA(b, cudaMemcpyAsync(b->dstate, h, sizeof(SolverState), cudaMemcpyHostToDevice, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  SolverState *st = VCL_DSTATE(b);

  const int grid = (int)std::max(1LL, std::min((n / 2 + VEC_THREADS - 1) / VEC_THREADS, (long long)std::min(b->sm_count * 8, VCL_MAX_BLOCKS)));
  const int kBatch = 32;
  int launched = 0;

  if (A->p2p)
  {
    // ---- peer-memory transport: 3 launches per iteration on one stream, scalars advanced by the SpMV kernel's last CTA ----
    double *loc_rr = b->dscal + 32;
    const u64 halo_base = A->halo_seq, red_base = A->red_seq;
    while (launched < tag->max_iterations)
    {
      const int nb = std::min(kBatch, tag->max_iterations - launched);
      VCL_RANGE("vcl:batch");
      for (int k = 0; k < nb; ++k)
      {
        // iteration i = launched + k + 1 uses exchange numbers base + i; once st->done is set every later kernel returns
        // at once on every rank (same sums -> same decision), so the exchanges executed are exactly 1..iters
        const u64 hseq = halo_base + (u64)(launched + k + 1), rseq = red_base + (u64)(launched + k + 1);
        PushRanges pr = PushRanges();
        if (A->fused_push)
        {
          const int par = (int)(hseq & 1ULL);
          pr.n = A->push.ndst; pr.me = A->push.me; pr.W = A->push.W; pr.seq = hseq;
          for (int d = 0; d < pr.n; ++d)
          {
            pr.lo[d] = A->push_lo[d]; pr.hi[d] = A->push_hi[d];
            pr.dst[d] = A->push.dst[d] + (size_t)par * (size_t)A->push.stride[d];
            pr.flag[d] = A->push.flag[d] + par * pr.W + pr.me;
          }
        }
        if (pr.n > 0) cg_update_push_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, x, p, r, Ap, 0.0, 0.0, st, b->partials, b->tickets, loc_rr, pr);
        else          cg_update_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, x, p, r, Ap, 0.0, 0.0, st, b->partials, b->tickets, loc_rr);
        VCL_LAUNCHED(b, "cg_update_kernel");
        // scattered send lists: the product kernel pushes from its own head (after its st->done test) -- no push kernel either way
        EpiFused<STEP_NONE, false, false> e = {Ap, p, nullptr, nullptr, b->partials, b->tickets, st, nullptr, nullptr, nullptr,
                                               {0.0, 0.0, 0.0}, nullptr, A->d_win, rseq, loc_rr, DIST_CG};
        if (A->fmt == 1) VCL_TRY(p2p_launch(b, A, hseq, p, e, !A->fused_push));
        else if (use_g)
        {
          CsrDev dd = p2p_all_blocks(b, A, hseq, true);
          dd.push = A->push_g.ndst > 0 ? A->d_push_g : nullptr;
          VCL_TRY((p2p_launch_csr<EpiFused<STEP_NONE, false, false>, false>(b, dd, make_xvec(p, 0, 1), e)));
        }
        else
        {
          CsrDev dd = p2p_all_blocks(b, A, hseq, !A->fused_push);
#ifdef VCL_PEER_DEBUG
          dd.dbg = A->hwin.dbg; dd.dbg_seq = rseq;
#endif
          VCL_TRY(p2p_launch_csr(b, dd, p2p_xvec(A, p, hseq), e));
        }
      }
      launched += nb;
      VCL_CUDA(b, cudaMemcpyAsync(h, st, sizeof(SolverState), cudaMemcpyDeviceToHost, b->stream));
      VCL_CUDA(b, cudaStreamSynchronize(b->stream));
      VCL_TRY(p2p_check(b, A));
      if (h->done != VCL_RUNNING) break;
    }
    A->halo_seq = halo_base + (u64)h->iters;
    A->red_seq = red_base + (u64)h->iters;
#ifdef VCL_PEER_DEBUG
    {
      std::vector<u64> d(4096);
      cudaMemcpy(d.data(), A->hwin.dbg, sizeof(u64) * 4096, cudaMemcpyDeviceToHost);
      double s_kernel = 0, s_red = 0, s_wait = 0; int cnt = 0;
      for (u64 q = red_base + 1; q <= red_base + (u64)h->iters && cnt < 1000; ++q, ++cnt)
      {
        const u64 *e = &d[(q % 1024) * 4];
        s_kernel += (double)(e[0] - e[2]); s_red += (double)(e[1] - e[0]); s_wait += (double)e[3];
      }
      if (cnt) fprintf(stderr, "[rank %d] peer debug over %d iterations: kernel start -> last CTA enters reduction %.1f us, reduction (push + wait + sum) %.1f us, max halo-flag wait per launch %.1f us\n",
                       b->rank, cnt, s_kernel / cnt * 1e-3, s_red / cnt * 1e-3, s_wait / cnt * 1e-3);
      cudaMemset(A->hwin.dbg, 0, sizeof(u64) * 4096);
    }
#endif
    tag->iters = h->iters;
    tag->error = std::sqrt(std::fabs(h->sums[0]) / norm_rhs_squared);
    return ViennaCLSuccess;
  }

  // rank-local sums land in `loc`, the allreduce writes the global sums into st->sums (out of place, so that re-issuing the
  // collective after convergence -- kernels skipped, `loc` unchanged -- reproduces the same global sums)
  double *loc = b->world > 1 ? b->dscal + 32 : &st->sums[0];
  if (b->world > 1) VCL_CUDA(b, cudaMemsetAsync(loc, 0, 3 * sizeof(double), b->stream));
  const NcclApi *api = b->world > 1 ? vcl_nccl(nullptr) : nullptr;
  while (launched < tag->max_iterations)
  {
    const int nb = std::min(kBatch, tag->max_iterations - launched);
    VCL_RANGE("vcl:batch");
    for (int k = 0; k < nb; ++k)
    {
      // NB: the kernels below are skipped on the device once st->done is set, but the collectives are still issued --
      // every rank takes the same decision from the same allreduced sums, so the call sequences stay matched.
      if (n > 0)
      {
        cg_update_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, x, p, r, Ap, 0.0, 0.0, st, b->partials, b->tickets, loc + 0);
        VCL_LAUNCHED(b, "cg_update_kernel");
      }
      VCL_TRY((nccl_fused_prod<false, false>(b, A, p, Ap, nullptr, nullptr, st, loc + 1, loc + 2, nullptr)));
      if (b->world > 1)
        VCL_NCCL(b, api, api->AllReduce(loc, &st->sums[0], 3, ncclDouble, ncclSum, (ncclComm_t)b->nccl_comm, b->stream));
      cg_advance_kernel<<<1, 1, 0, b->stream>>>(st);
      VCL_LAUNCHED(b, "cg_advance_kernel");
    }
    launched += nb;
    VCL_CUDA(b, cudaMemcpyAsync(h, st, sizeof(SolverState), cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    VCL_TRY(vcl_comm_check(b));                            // asynchronous NCCL errors (a peer died, a transport failed) surface here
    if (h->done != VCL_RUNNING) break;
  }
  tag->iters = h->iters;
  tag->error = std::sqrt(std::fabs(h->sums[0]) / norm_rhs_squared);
  return ViennaCLSuccess;
}

// bicgstab.hpp:97-215 (pipelined BiCGStab, no preconditioner) over row-partitioned data.  Per iteration: two fused products, each
// with its own halo exchange and its own all-reduce in the product kernel's tail ({<r,r0*>, <Ap,r0*>} after Ap = A p;
// {<s,s>, <As,As>, <As,s>, <As,r0*>} after As = A s, followed by the scalar recurrences), and two vector kernels.
ViennaCLStatus ViennaCLCUDADdist_csr_bicgstab(ViennaCLBackend b, ViennaCLB200DistCsr A, const double *rhs, double *x, ViennaCLB200SolverTag *tag)
{
  VCL_RANGE("vcl:dist_bicgstab");
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, A && tag, "bad arguments");
  VCL_REQUIRE(b, tag->precond == ViennaCLB200PrecondNone, "row-partitioned BiCGStab: only the unpreconditioned pipelined path is provided");
  VCL_REQUIRE(b, tag->monitor == nullptr, "monitor callbacks are not supported on the row-partitioned path");
  const long long n = A->n;
  tag->iters = 0; tag->error = 0.0;
  VCL_REQUIRE(b, n == 0 || (rhs && x), "null vector");
  VCL_CUDA(b, cudaSetDevice(b->device));
  const size_t need = ((size_t)std::max<long long>(n, 1) * sizeof(double) + 255) / 256 * 256;
  VCL_TRY(vcl_ws_reserve(b, 6 * need));
  char *w0 = (char*)b->ws;
  double *r = (double*)w0, *p = (double*)(w0 + need), *r0 = (double*)(w0 + 2 * need), *Ap = (double*)(w0 + 3 * need), *s = (double*)(w0 + 4 * need),
         *As = (double*)(w0 + 5 * need);
  VCL_CUDA(b, cudaMemsetAsync(x, 0, sizeof(double) * n, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(r, rhs, sizeof(double) * n, cudaMemcpyDeviceToDevice, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(p, rhs, sizeof(double) * n, cudaMemcpyDeviceToDevice, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(r0, rhs, sizeof(double) * n, cudaMemcpyDeviceToDevice, b->stream));
  // rank-local totals, laid out like SolverState::sums: [0] <r,r0*>, [1] <As,As>, [2] <As,s>, [3] <Ap,r0*>, [4] <As,r0*>, [5] <s,s>
  double *loc = b->dscal + 32;
  VCL_CUDA(b, cudaMemsetAsync(loc, 0, 8 * sizeof(double), b->stream));
  if (n > 0) VCL_TRY(vcl_dot_async(b, n, r, 0, 1, r, 0, 1, loc + 0));
  VCL_CUDA(b, cudaMemcpyAsync(b->dscal, loc, sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
  VCL_TRY(allreduce_sum(b, A, b->dscal, 1));
  VCL_CUDA(b, cudaMemcpyAsync(b->hscal, b->dscal, sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  const double norm_rhs = std::sqrt(b->hscal[0]);
  if (norm_rhs <= tag->abs_tolerance) return ViennaCLSuccess;                       // bicgstab.hpp:140-141

  SolverState *h = VCL_HSTATE(b);
  std::memset(h, 0, sizeof(SolverState));
  h->norm_rhs = norm_rhs; h->norm_rhs_sq = norm_rhs * norm_rhs; h->residual_norm = norm_rhs;
  h->tol = tag->tolerance; h->abs_tol = tag->abs_tolerance; h->maxit = tag->max_iterations;
  h->sums[0] = norm_rhs * norm_rhs;                                                 // bicgstab.hpp:131
  VCL_CUDA(b, cudaMemcpyAsync(b->dstate, h, sizeof(SolverState), cudaMemcpyHostToDevice, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  SolverState *st = VCL_DSTATE(b);

  const int grid = (int)std::max(1LL, std::min((n / 2 + VEC_THREADS - 1) / VEC_THREADS, (long long)std::min(b->sm_count * 8, VCL_MAX_BLOCKS)));
  const int kBatch = 32;
  int launched = 0;
  const u64 halo_base = A->halo_seq, red_base = A->red_seq;
  while (launched < tag->max_iterations)
  {
    const int nb = std::min(kBatch, tag->max_iterations - launched);
    VCL_RANGE("vcl:batch");
    for (int k = 0; k < nb; ++k)
    {
      const u64 it = (u64)(launched + k);                  // iteration it + 1 uses exchanges 2 it + 1 and 2 it + 2
      if (A->p2p)
      {
        EpiFused<STEP_NONE, true, false> e1 = {Ap, p, r0, nullptr, b->partials, b->tickets, st, nullptr, nullptr, nullptr,
                                               {0.0, 0.0, 0.0}, nullptr, A->d_win, red_base + 2 * it + 1, loc + 0, DIST_BICG_P};
        VCL_TRY(p2p_launch(b, A, halo_base + 2 * it + 1, p, e1, true));
      }
      else
      {
        VCL_TRY((nccl_fused_prod<true, false>(b, A, p, Ap, r0, nullptr, st, nullptr, nullptr, loc + 3)));
        VCL_TRY(nccl_sums(b, loc, &st->sums[0], 4));      // <r,r0*> and <Ap,r0*> are needed now (alpha); [1], [2] are rewritten below
      }
      if (n > 0)
      {
        bicgstab_update_s_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, s, r, Ap, &st->sums[0], &st->sums[3], st, b->partials, b->tickets, loc + 5);
        VCL_LAUNCHED(b, "bicgstab_update_s_kernel");
      }
      if (A->p2p)
      {
        EpiFused<STEP_NONE, true, false> e2 = {As, s, r0, nullptr, b->partials, b->tickets, st, nullptr, nullptr, nullptr,
                                               {0.0, 0.0, 0.0}, nullptr, A->d_win, red_base + 2 * it + 2, loc + 5, DIST_BICG_S};
        VCL_TRY(p2p_launch(b, A, halo_base + 2 * it + 2, s, e2, true));
      }
      else
      {
        VCL_TRY((nccl_fused_prod<true, false>(b, A, s, As, r0, nullptr, st, loc + 1, loc + 2, loc + 4)));
        VCL_TRY(nccl_sums(b, loc, &st->sums[0], 6));
        bicgstab_advance_kernel<<<1, 1, 0, b->stream>>>(st);
        VCL_LAUNCHED(b, "bicgstab_advance_kernel");
      }
      if (n > 0)
      {
        bicgstab_update_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, x, 0.0, p, 0.0, s, r, As, 0.0, Ap, r0, st, b->partials, b->tickets, loc + 0);
        VCL_LAUNCHED(b, "bicgstab_update_kernel");
      }
      else
      {
        // a rank without rows still has to raise MAXIT like the others (bicgstab_update_kernel does it for them)
        bicgstab_maxit_kernel<<<1, 1, 0, b->stream>>>(st);
        VCL_LAUNCHED(b, "bicgstab_maxit_kernel");
      }
    }
    launched += nb;
    VCL_TRY(dist_pull_state(b, A));
    if (h->done != VCL_RUNNING) break;
  }
  if (A->p2p) { A->halo_seq = halo_base + 2 * (u64)h->iters; A->red_seq = red_base + 2 * (u64)h->iters; }
  tag->iters = h->iters;
  tag->error = h->residual_norm / norm_rhs;                                          // bicgstab.hpp:212
  return ViennaCLSuccess;
}

// gmres.hpp:181-367 (pipelined GMRES(m), classical Gram-Schmidt, no preconditioner) over row-partitioned data.  The basis lives
// in slabs; per inner iteration: one row-partitioned product (halo exchange inside the launch on the peer-memory transport),
// Gram-Schmidt stage 1 on the slab -> all-reduce of the k dots -> stage 2 -> all-reduce of ||v_k||^2 -> normalise; the
// <r, v_k> of a cycle are all-reduced once, at its end.  R and xi are then global, so the host half of the cycle
// (vcl_gmres_cycle_host) takes identical decisions on every rank.
ViennaCLStatus ViennaCLCUDADdist_csr_gmres(ViennaCLBackend b, ViennaCLB200DistCsr A, const double *rhs, double *x, ViennaCLB200SolverTag *tag)
{
  VCL_RANGE("vcl:dist_gmres");
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, A && tag, "bad arguments");
  VCL_REQUIRE(b, tag->precond == ViennaCLB200PrecondNone, "row-partitioned GMRES: only the unpreconditioned pipelined path is provided");
  VCL_REQUIRE(b, tag->monitor == nullptr, "monitor callbacks are not supported on the row-partitioned path");
  VCL_REQUIRE(b, tag->krylov_dim >= 1 && tag->krylov_dim <= VCL_GMRES_MAX_KRYLOV, "krylov_dim must be in [1, 64]");
  const long long n = A->n;
  tag->iters = 0; tag->error = 0.0;
  VCL_REQUIRE(b, n == 0 || (rhs && x), "null vector");
  VCL_CUDA(b, cudaSetDevice(b->device));
  const int m = tag->krylov_dim;
  const long long isz = (std::max<long long>(n, 1) + 127) / 128 * 128;
  auto need = [](size_t cnt) { return (cnt * sizeof(double) + 255) / 256 * 256; };
  VCL_TRY(vcl_ws_reserve(b, need((size_t)std::max<long long>(n, 1)) + need((size_t)isz * m) + need((size_t)m * m) + 3 * need(m)));
  char *w0 = (char*)b->ws;
  double *res = (double*)w0;               w0 += need((size_t)std::max<long long>(n, 1));
  double *V = (double*)w0;                 w0 += need((size_t)isz * m);
  double *R = (double*)w0;                 w0 += need((size_t)m * m);
  double *d_xi = (double*)w0;              w0 += need(m);
  double *d_h = (double*)w0;               w0 += need(m);
  double *d_coef = (double*)w0;
  double *d_nsq = b->dscal + 8, *d_ss = b->dscal + 9;
  std::vector<double> hR((size_t)m * m), xi(m), eta(m), coef(m, 0.0);

  VCL_CUDA(b, cudaMemsetAsync(x, 0, sizeof(double) * n, b->stream));
  VCL_CUDA(b, cudaMemcpyAsync(res, rhs, sizeof(double) * n, cudaMemcpyDeviceToDevice, b->stream));
  VCL_CUDA(b, cudaMemsetAsync(R, 0, sizeof(double) * m * m, b->stream));
  VCL_CUDA(b, cudaMemsetAsync(d_xi, 0, sizeof(double) * m, b->stream));
  VCL_CUDA(b, cudaMemsetAsync(d_coef, 0, sizeof(double) * m, b->stream));
  // global <v, v> of a slab vector, on the host
  auto global_ss = [&](const double *v, double *out) -> ViennaCLStatus
  {
    VCL_CUDA(b, cudaMemsetAsync(d_ss, 0, sizeof(double), b->stream));
    if (n > 0) VCL_TRY(vcl_dot_async(b, n, v, 0, 1, v, 0, 1, d_ss));
    VCL_TRY(allreduce_sum(b, A, d_ss, 1));
    VCL_CUDA(b, cudaMemcpyAsync(b->hscal, d_ss, sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    *out = b->hscal[0];
    return ViennaCLSuccess;
  };
  double ss = 0.0;
  VCL_TRY(global_ss(res, &ss));
  const double norm_rhs = std::sqrt(ss);
  double rho_0 = norm_rhs, rho = 1.0;

  unsigned max_restarts = (unsigned)tag->max_iterations / (unsigned)m;         // gmres.hpp:74-80
  if (max_restarts > 0 && max_restarts * (unsigned)m == (unsigned)tag->max_iterations) max_restarts -= 1;
  const int grid = scalar_grid(b, std::max<long long>(n, 1));

  for (unsigned restart = 0; restart <= max_restarts; ++restart)
  {
    VCL_RANGE("vcl:gmres_cycle");
    if (restart > 0)
    {
      VCL_TRY(dist_plain_prod(b, A, x, res));
      if (n > 0) { residual_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, res, rhs); VCL_LAUNCHED(b, "residual_kernel"); }
      VCL_TRY(global_ss(res, &ss));
      rho_0 = std::sqrt(ss);
    }
    if (rho_0 <= tag->abs_tolerance) break;                                     // gmres.hpp:227-228
    if (n > 0) { scale_residual_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, res, rho_0); VCL_LAUNCHED(b, "scale_residual_kernel"); }
    rho = 1.0;
    if (rho_0 / norm_rhs < tag->tolerance || rho_0 < tag->abs_tolerance) break;  // gmres.hpp:234-235

    for (int k = 0; k < m; ++k)
    {
      double *vk = V + (size_t)k * isz;
      const double *src = (k == 0) ? res : V + (size_t)(k - 1) * isz;
      VCL_TRY(dist_plain_prod(b, A, src, vk));
      VCL_CUDA(b, cudaMemsetAsync(d_nsq, 0, sizeof(double), b->stream));
      if (k > 0)
      {
        VCL_CUDA(b, cudaMemsetAsync(d_h, 0, sizeof(double) * k, b->stream));
        if (n > 0) VCL_TRY(launch_gs1(b, grid, V, n, isz, k, d_h, 1));
        VCL_TRY(allreduce_sum(b, A, d_h, k));
        if (n > 0)
        {
          gmres_gs2_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(V, n, isz, k, d_h, 1, R, m, d_nsq, b->partials, b->tickets);
          VCL_LAUNCHED(b, "gmres_gs2_kernel");
        }
        else                                                 // a rank without rows still keeps its copy of R (the host half reads it)
          VCL_CUDA(b, cudaMemcpy2DAsync(R + (size_t)k * m, sizeof(double), d_h, sizeof(double), sizeof(double), k, cudaMemcpyDeviceToDevice, b->stream));
      }
      else if (n > 0) VCL_TRY(vcl_dot_async(b, n, vk, 0, 1, vk, 0, 1, d_nsq));
      VCL_TRY(allreduce_sum(b, A, d_nsq, 1));
      gmres_normalize_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, vk, res, R, k * m + k, d_nsq, d_xi + k, b->partials, b->tickets);
      VCL_LAUNCHED(b, "gmres_normalize_kernel");
    }
    VCL_TRY(allreduce_sum(b, A, d_xi, m));
    VCL_CUDA(b, cudaMemcpyAsync(xi.data(), d_xi, sizeof(double) * m, cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaMemcpyAsync(hR.data(), R, sizeof(double) * m * m, cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    if (A->p2p) VCL_TRY(p2p_check(b, A)); else VCL_TRY(vcl_comm_check(b));

    bool converged = false;
    const size_t kk = vcl_gmres_cycle_host<double>(tag, m, hR, xi, eta, coef, rho, rho_0, norm_rhs, false, &converged);
    VCL_CUDA(b, cudaMemcpyAsync(d_coef, coef.data(), sizeof(double) * m, cudaMemcpyHostToDevice, b->stream));
    if (n > 0) { gmres_update_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, x, res, V, isz, d_coef, (int)kk); VCL_LAUNCHED(b, "gmres_update_kernel"); }
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));                               // coef (pageable) must outlive the copy
    tag->error = std::fabs(rho * rho_0 / norm_rhs);                              // gmres.hpp:360
  }
  return ViennaCLSuccess;
}

} // extern "C"
