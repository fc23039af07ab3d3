// dist.cu -- row-partitioned CSR SpMV and pipelined CG across the GPUs of one box, one process per GPU.
// No counterpart in the reference ("Partition of data is left to the user", doc/manual/multi-device.dox:9).
//
// Rank g owns a contiguous block of rows.  At create time the locally stored (global) column indices are analysed on the
// device: columns outside the owned range become the halo; they are renumbered to [owned | halo] (halo sorted by global id,
// hence grouped by owner), every owner learns which of its entries each peer needs (send lists), and the CSR row blocks
// are split into INTERIOR blocks (no halo column) and BOUNDARY blocks.
// One product:   pack owned entries -> ncclSend/ncclRecv with the peers on the communication stream
//                || interior row blocks on the compute stream            (overlap)
//                then boundary row blocks once the halo has landed.
// One CG iteration adds a single 3-double allreduce (<r,r>, <Ap,Ap>, <p,Ap>) -- the property Chronopoulos/Gear CG was
// chosen for (cg.hpp:116-118) -- after which a one-thread kernel advances alpha/beta/convergence on every rank identically.
#include "fused_kernels.cuh"
#include "launch.cuh"
#include "blas1.cuh"
#include "nccl_dyn.cuh"
#include <cmath>
#include <algorithm>

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
  u32 *interior = nullptr, *boundary = nullptr; int n_interior = 0, n_boundary = 0;
  double *tmp_sums = nullptr;        // 3 doubles: interior totals
  cudaEvent_t ev_x = nullptr, ev_halo = nullptr;
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

__global__ void cg_advance_kernel(SolverState *st)
{
  if (st->done == VCL_RUNNING) cg_advance(st);
}

ViennaCLStatus build_row_blocks_host(const std::vector<u32> &rp, int rows, std::vector<u32> &blk)
{
  blk.clear(); blk.push_back(0);
  int r = 0;
  while (r < rows)
  {
    const u32 start = rp[r], slack = start & 3u;
    int e = r;
    while (e < rows && e - r < VCL_B200_CSR_BLOCK_ROWS && (rp[e + 1] - start) + slack <= VCL_B200_CSR_BLOCK_NNZ) ++e;
    if (e == r) e = r + 1;
    blk.push_back((u32)e);
    r = e;
  }
  return ViennaCLSuccess;
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

CsrDev subset(ViennaCLB200DistCsr A, bool boundary)
{
  CsrDev d = {A->n, (u32)A->nnz, A->rp, A->ci_local, A->va, A->blk, boundary ? A->n_boundary : A->n_interior,
              boundary ? A->boundary : A->interior};
  return d;
}

ViennaCLStatus allreduce_sum(ViennaCLBackend b, double *buf, int count)
{
  if (b->world == 1) return ViennaCLSuccess;
  const NcclApi *api = vcl_nccl(nullptr);
  VCL_NCCL(b, api, api->AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)b->nccl_comm, b->stream));
  return ViennaCLSuccess;
}

// y = A x with overlap; epilogue factory gives the interior / boundary epilogues
ViennaCLStatus dist_plain_prod(ViennaCLBackend b, ViennaCLB200DistCsr A, const double *x, double *y)
{
  XVec xv = {x, 0, 1, A->halo_buf, (u32)A->n};
  EpiAxpby epi = {y, 0, 1, 1.0, 0.0};
  VCL_TRY(start_halo(b, A, x));
  VCL_TRY(vcl_launch_csr_split(b, subset(A, false), xv, epi, b->stream));
  VCL_TRY(wait_halo(b, A));
  VCL_TRY(vcl_launch_csr_split(b, subset(A, true), xv, epi, b->stream));
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
  if (W > 1)
  {
    int *d_cnt = nullptr;
    VCL_CUDA(b, cudaMalloc(&d_cnt, sizeof(int) * (size_t)W * (W + 1)));
    VCL_CUDA(b, cudaMemcpyAsync(d_cnt + (size_t)W * W, A->recv_cnt.data(), sizeof(int) * W, cudaMemcpyHostToDevice, b->stream));
    VCL_NCCL(b, api, api->AllGather(d_cnt + (size_t)W * W, d_cnt, (size_t)W, ncclInt32, comm, b->stream));
    std::vector<int> M((size_t)W * W);
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
    std::vector<u32> rp((size_t)A->n + 1), blk;
    VCL_CUDA(b, cudaMemcpyAsync(rp.data(), row_ptr, sizeof(u32) * ((size_t)A->n + 1), cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    build_row_blocks_host(rp, A->n, blk);
    A->nblk = (int)blk.size() - 1;
    VCL_CUDA(b, cudaMalloc(&A->blk, sizeof(u32) * blk.size()));
    VCL_CUDA(b, cudaMemcpyAsync(A->blk, blk.data(), sizeof(u32) * blk.size(), cudaMemcpyHostToDevice, b->stream));
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
    VCL_CUDA(b, cudaMalloc(&A->interior, sizeof(u32) * std::max<size_t>(in.size(), 1)));
    VCL_CUDA(b, cudaMalloc(&A->boundary, sizeof(u32) * std::max<size_t>(bd.size(), 1)));
    if (!in.empty()) VCL_CUDA(b, cudaMemcpyAsync(A->interior, in.data(), sizeof(u32) * in.size(), cudaMemcpyHostToDevice, b->stream));
    if (!bd.empty()) VCL_CUDA(b, cudaMemcpyAsync(A->boundary, bd.data(), sizeof(u32) * bd.size(), cudaMemcpyHostToDevice, b->stream));
  }
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  if (d_halo) VCL_CUDA(b, cudaFree(d_halo));
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
  cudaFree(A->send_idx); cudaFree(A->send_buf); cudaFree(A->halo_buf); cudaFree(A->blk); cudaFree(A->interior); cudaFree(A->boundary);
  cudaFree(A->tmp_sums);
  if (A->ev_x) cudaEventDestroy(A->ev_x);
  if (A->ev_halo) cudaEventDestroy(A->ev_halo);
  delete A;
  *pA = nullptr;
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDADdist_csrmv(ViennaCLBackend b, ViennaCLB200DistCsr A, const double *x_local, double *y_local)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, A && (A->n == 0 || (x_local && y_local)) && x_local != y_local, "bad arguments");
  if (A->n == 0 && b->world == 1) return ViennaCLSuccess;
  return dist_plain_prod(b, A, x_local, y_local);
}

// cg.hpp:128-187 over row-partitioned data: local fused kernels + halo exchange + one allreduce per iteration.
ViennaCLStatus ViennaCLCUDADdist_csr_cg(ViennaCLBackend b, ViennaCLB200DistCsr A, const double *rhs, double *x, ViennaCLB200SolverTag *tag)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, A && tag, "bad arguments");
  VCL_REQUIRE(b, tag->precond == ViennaCLB200PrecondNone, "CG: only the unpreconditioned pipelined path is provided");
  VCL_REQUIRE(b, tag->monitor == nullptr, "monitor callbacks are not supported on the row-partitioned path");
  const long long n = A->n;
  tag->iters = 0; tag->error = 0.0;
  VCL_REQUIRE(b, n == 0 || (rhs && x), "null vector");
  VCL_CUDA(b, cudaSetDevice(b->device));
  const size_t need = ((size_t)std::max<long long>(n, 1) * sizeof(double) + 255) / 256 * 256;
  VCL_TRY(vcl_ws_reserve(b, 3 * need));
  double *r = (double*)b->ws, *p = (double*)((char*)b->ws + need), *Ap = (double*)((char*)b->ws + 2 * need);

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
  VCL_TRY(allreduce_sum(b, b->dscal, 3));
  VCL_CUDA(b, cudaMemcpyAsync(b->hscal, b->dscal, 3 * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));

  double norm_rhs_squared = std::sqrt(b->hscal[0]); norm_rhs_squared *= norm_rhs_squared;
  if (norm_rhs_squared <= tag->abs_tolerance * tag->abs_tolerance) return ViennaCLSuccess;
  const double rr = norm_rhs_squared;
  const double alpha = rr / b->hscal[1];
  double beta = std::sqrt(b->hscal[2]); beta = (alpha * alpha * beta * beta - rr) / rr;

  SolverState *h = b->hstate;
  std::memset(h, 0, sizeof(SolverState));
  h->alpha = alpha; h->beta = beta; h->norm_rhs_sq = norm_rhs_squared; h->norm_rhs = std::sqrt(norm_rhs_squared);
  h->tol = tag->tolerance; h->abs_tol = tag->abs_tolerance; h->maxit = tag->max_iterations; h->sums[0] = rr;
  VCL_CUDA(b, cudaMemcpyAsync(b->dstate, h, sizeof(SolverState), cudaMemcpyHostToDevice, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  SolverState *st = b->dstate;

  const int grid = (int)std::max(1LL, std::min((n / 2 + VEC_THREADS - 1) / VEC_THREADS, (long long)std::min(b->sm_count * 8, VCL_MAX_BLOCKS)));
  XVec xv = {p, 0, 1, A->halo_buf, (u32)A->n};
  // rank-local sums land in `loc`, the allreduce writes the global sums into st->sums (out of place, so that re-issuing the
  // collective after convergence -- kernels skipped, `loc` unchanged -- reproduces the same global sums)
  double *loc = b->world > 1 ? b->dscal + 32 : &st->sums[0];
  if (b->world > 1) VCL_CUDA(b, cudaMemsetAsync(loc, 0, 3 * sizeof(double), b->stream));
  const NcclApi *api = b->world > 1 ? vcl_nccl(nullptr) : nullptr;
  const int kBatch = 32;
  int launched = 0;
  while (launched < tag->max_iterations)
  {
    const int nb = std::min(kBatch, tag->max_iterations - launched);
    for (int k = 0; k < nb; ++k)
    {
      // NB: the kernels below are skipped on the device once st->done is set, but the collectives are still issued --
      // every rank takes the same decision from the same allreduced sums, so the call sequences stay matched.
      if (n > 0)
      {
        cg_update_kernel<<<grid, VEC_THREADS, 0, b->stream>>>(n, x, p, r, Ap, 0.0, 0.0, st, b->partials, b->tickets, loc + 0);
        VCL_LAUNCHED(b, "cg_update_kernel");
      }
      VCL_TRY(start_halo(b, A, p));
      if (A->n_interior > 0)
      {
        EpiFused<STEP_NONE, false, false> ei = {Ap, p, nullptr, nullptr, b->partials, b->tickets, st,
                                                A->n_boundary > 0 ? A->tmp_sums + 0 : loc + 1,
                                                A->n_boundary > 0 ? A->tmp_sums + 1 : loc + 2, nullptr, {0.0, 0.0, 0.0}, nullptr};
        VCL_TRY(vcl_launch_csr_split(b, subset(A, false), xv, ei, b->stream));
      }
      VCL_TRY(wait_halo(b, A));
      if (A->n_boundary > 0)
      {
        EpiFused<STEP_NONE, false, false> eb = {Ap, p, nullptr, nullptr, b->partials, b->tickets, st, loc + 1, loc + 2, nullptr,
                                                {0.0, 0.0, 0.0}, A->n_interior > 0 ? A->tmp_sums : nullptr};
        VCL_TRY(vcl_launch_csr_split(b, subset(A, true), xv, eb, b->stream));
      }
      if (b->world > 1)
        VCL_NCCL(b, api, api->AllReduce(loc, &st->sums[0], 3, ncclDouble, ncclSum, (ncclComm_t)b->nccl_comm, b->stream));
      cg_advance_kernel<<<1, 1, 0, b->stream>>>(st);
      VCL_LAUNCHED(b, "cg_advance_kernel");
    }
    launched += nb;
    VCL_CUDA(b, cudaMemcpyAsync(h, st, sizeof(SolverState), cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    if (h->done != VCL_RUNNING) break;
  }
  tag->iters = h->iters;
  tag->error = std::sqrt(std::fabs(h->sums[0]) / norm_rhs_squared);
  return ViennaCLSuccess;
}

} // extern "C"
