// peer.cuh -- peer-memory (NVLink / NVSwitch) primitives of the row-partitioned path: every rank maps a small "window" of
// every other rank's device memory (CUDA IPC) and the kernels of the solver push halo entries and reduction partials
// straight into the consumer's memory, followed by a release-store of a sequence number; consumers spin on their OWN
// memory with acquire loads.  No communication kernel, no second stream, no host involvement inside an iteration:
//   halo_push_share()    called at the head of the product launch: its first CTAs send this rank's boundary x entries to the
//                        neighbours' halo receive buffers and publish the flag                 (no launch of its own)
//   csr_stream_kernel /  interior row blocks (passes) first; boundary blocks wait for the flags (the SpMV launch itself)
//   sell_kernel
//   peer_allreduce()     called by the LAST CTA of the reducing kernel: push the rank-local totals to all ranks, wait
//                        for everyone's, sum in rank order -> bit-identical global sums on every rank
// Buffers are double-buffered by the parity of the sequence number; every exchange pair is symmetric (a rank that sends
// to q also waits for q), which bounds any rank to at most one exchange ahead of its partners.
// No counterpart in the reference (doc/manual/multi-device.dox:9).
#pragma once
#include "common.cuh"

#define VCL_MAX_PEERS 16
#define VCL_PEER_TIMEOUT_NS 8000000000ULL      // a spin that lasts 8 s is reported (error word) instead of hanging the GPU

typedef unsigned long long u64;

// All pointers are valid in THIS process (own window: the allocation itself; peers: cudaIpcOpenMemHandle mappings).
struct PeerWindow
{
  int W, me;
  double *halo[VCL_MAX_PEERS];          // rank q: halo receive area  [2][halo_len[q]]
  long long halo_len[VCL_MAX_PEERS];
  u64 *halo_flag[VCL_MAX_PEERS];        // rank q: [2][W]  (indexed by source rank)
  double *red[VCL_MAX_PEERS];           // rank q: [2][W][4]
  u64 *red_flag[VCL_MAX_PEERS];         // rank q: [2][W]
  int *err;                             // own error word (set on time-out)
  u64 *dbg;                             // VCL_PEER_DEBUG builds: [1024][4] time stamps (kernel start, last CTA enters the
                                        // reduction, reduction done) -- NULL otherwise
};

// One destination of a halo push (passed by value to the kernel).
struct HaloPush
{
  int ndst, me, W;
  int begin[VCL_MAX_PEERS + 1];         // segment of the send list per destination
  double *dst[VCL_MAX_PEERS];           // destination's halo area, already offset to this rank's segment (parity 0)
  long long stride[VCL_MAX_PEERS];      // destination's halo_len (parity stride)
  u64 *flag[VCL_MAX_PEERS];             // destination's halo_flag (parity 0), entry of this rank added in the kernel
};

// Halo push fused into the kernel that PRODUCES the vector (cg_update_kernel): usable when every destination's send list
// is one contiguous index range [lo, hi) -- slab partitions of banded / stencil matrices.  n == 0: nothing to push.
#define VCL_MAX_PUSH_RANGES 4
struct PushRanges
{
  int n, me, W;
  long long lo[VCL_MAX_PUSH_RANGES], hi[VCL_MAX_PUSH_RANGES];
  double *dst[VCL_MAX_PUSH_RANGES];     // destination's halo area, offset to this rank's segment AND to the parity buffer
  u64 *flag[VCL_MAX_PUSH_RANGES];       // destination's halo_flag entry for (parity, this rank)
  u64 seq;
};

#ifdef __CUDACC__
__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p)
{
  u64 v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(u64 *p, u64 v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 global_ns()
{
  u64 t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Error word raised on a time-out: kind (1: halo flag, 2: reduction flag) in bits 28.., the partner rank in bits 20..27 and the
// low 20 bits of the sequence number of the exchange that did not complete; the first time-out wins.  The results of the launch
// that timed out (and of everything enqueued after it) are undefined; the host reports the decoded word at the next batch
// boundary (dist.cu: p2p_check) -- "which exchange, waiting for whom".
#define VCL_PEER_ERR(kind, partner, seq) ((int)(((unsigned)(kind) << 28) | (((unsigned)(partner) & 0xffu) << 20) | ((unsigned)(seq) & 0xfffffu)))

// Spins until *flag >= seq (acquire).  Returns false on time-out after raising *err.
__device__ __forceinline__ bool peer_wait(const u64 *flag, u64 seq, int *err, int kind = 1, int partner = 0)
{
  if (ld_acquire_sys(flag) >= seq) return true;
  const u64 t0 = global_ns();
  for (;;)
  {
    __nanosleep(40);
    if (ld_acquire_sys(flag) >= seq) return true;
    if (global_ns() - t0 > VCL_PEER_TIMEOUT_NS) { if (err) atomicCAS(err, 0, VCL_PEER_ERR(kind, partner, seq)); return false; }
  }
}

// returns true when entry k went to at least one neighbour (the caller fences only then)
__device__ __forceinline__ bool push_entry(const PushRanges &pr, long long k, double v)
{
  bool sent = false;
#pragma unroll
  for (int d = 0; d < VCL_MAX_PUSH_RANGES; ++d)
    if (d < pr.n && k >= pr.lo[d] && k < pr.hi[d]) { pr.dst[d][k - pr.lo[d]] = v; sent = true; }
  return sent;
}

// Halo push from the head of the kernel that consumes the halo (the row-partitioned product is ONE launch): the first K CTAs
// of the grid -- the ones the hardware schedules first -- send x[idx[i]] to the destinations' receive buffers, ~4 entries
// per thread; every thread fences its own remote stores and the CTA that takes the last of the K tickets publishes the
// sequence number to every destination with a release store.  Nothing waits here, and the other CTAs go straight to
// their row blocks.
template<class T>    // T = double; a template only because the single-precision build also compiles (never runs) the row-partitioned branch
__device__ __forceinline__ void halo_push_share(const HaloPush &hp, const unsigned int *idx, const T *x, u64 seq, unsigned int *ticket)
{
  __shared__ bool s_push_last;
  const int ndst = hp.ndst, total = hp.begin[ndst];
  const unsigned int K = min(gridDim.x, (unsigned int)max(1, (total + 1023) / 1024));
  if (blockIdx.x >= K) return;                              // uniform over the CTA
  const int par = (int)(seq & 1ULL);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += K * blockDim.x)
  {
    int d = 0;
    while (i >= hp.begin[d + 1]) ++d;
    hp.dst[d][(size_t)par * hp.stride[d] + (size_t)(i - hp.begin[d])] = x[idx[i]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_push_last = (atomicAdd(ticket, 1u) == K - 1);
  __syncthreads();
  if (s_push_last)
  {
    __threadfence_system();
    if ((int)threadIdx.x < ndst) st_release_sys(hp.flag[threadIdx.x] + par * hp.W + hp.me, seq);
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

// All-reduce (sum) of N <= 4 doubles across the ranks, executed by ONE CTA per rank (>= W threads).
// vals: shared memory, N inputs (valid before the call, written by thread 0) -> N global sums (valid for thread 0 after).
// gather: shared memory, >= W*4 doubles.
template<int N>
__device__ __forceinline__ void peer_allreduce(const PeerWindow *win, u64 seq, double *vals, double *gather)
{
  const int W = win->W, me = win->me, par = (int)(seq & 1ULL), tid = threadIdx.x;
  __syncthreads();
  if (tid < W)
  {
    double *slot = win->red[tid] + (size_t)(par * W + me) * 4;
#pragma unroll
    for (int j = 0; j < N; ++j) __stcg(slot + j, vals[j]);
    __threadfence_system();
    st_release_sys(win->red_flag[tid] + par * W + me, seq);
    // every rank's contribution lands in my own window
    peer_wait(win->red_flag[me] + par * W + tid, seq, win->err, 2, tid);
    const double *mine = win->red[me] + (size_t)(par * W + tid) * 4;
#pragma unroll
    for (int j = 0; j < N; ++j) gather[tid * 4 + j] = __ldcg(mine + j);
  }
  __syncthreads();
  if (tid == 0)
  {
#pragma unroll
    for (int j = 0; j < N; ++j)
    {
      double s = 0.0;
      for (int q = 0; q < W; ++q) s += gather[q * 4 + j];      // rank order: identical on every rank
      vals[j] = s;
    }
  }
}
#endif
