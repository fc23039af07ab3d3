// spmv.cu -- C-ABI entry points for CSR / SELL SpMV, row blocks, row_info and CSR->SELL conversion.
#include "spmv_kernels.cuh"
#include "launch.cuh"
#include <cstdlib>

namespace VCL_NS
{

#ifndef VCL_F32      // index-only helpers: compiled once (double build)
// ------------------------------------------------------------------------------------------------
// Row blocks (compressed_matrix.hpp:1152-1188 analogue).  Greedy over whole rows: a block closes when the next row would
// exceed VCL_B200_CSR_BLOCK_NNZ staged entries or VCL_B200_CSR_BLOCK_ROWS rows; a longer row gets a block of its own.
// Like the reference this runs on the host after a D2H copy of row_ptr (set-up path, outside every timed region).
// ------------------------------------------------------------------------------------------------
static void build_row_blocks(const u32 *rp, int rows, std::vector<u32> &blk)
{
  blk.clear();
  blk.push_back(0);
  int r = 0;
  while (r < rows)
  {
    const u32 start = rp[r];
    // staged range begins at the 4-aligned entry below `start`
    const u32 slack = start & 3u;
    int e = r;
    while (e < rows && e - r < VCL_B200_CSR_BLOCK_ROWS && (rp[e + 1] - start) + slack <= VCL_B200_CSR_BLOCK_NNZ) ++e;
    if (e == r) e = r + 1;          // single long row
    blk.push_back((u32)e);
    r = e;
  }
}

// Device-side plan for matrices whose rows are short and evenly filled (every stencil / FEM matrix of the benchmarks):
// for G = 256, 128, ..., 1 the largest number of staged entries of any ALIGNED group of G rows is one difference of two
// row pointers per group; the largest G whose groups all fit VCL_B200_CSR_BLOCK_NNZ gives the plan blk[i] = min(i*G, rows)
// without copying row_ptr (4 bytes per row: 537 MB at 512^3) to the host.  Anything else (a row longer than a block, or
// groups that only fit at G < 32) takes the greedy host scan above, like the reference (compressed_matrix.hpp:1152-1188).
__global__ void row_group_fill_kernel(int rows, const u32 * __restrict__ rp, unsigned int *max_fill /* [9] */)
{
  __shared__ unsigned int s_max[9];
  if (threadIdx.x < 9) s_max[threadIdx.x] = 0u;
  __syncthreads();
  unsigned int loc[9] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x)
  {
    const u32 start = rp[r];
#pragma unroll
    for (int lg = 0; lg < 9; ++lg)                       // group size G = 1 << lg; row r opens a group when r % G == 0
    {
      const long long G = 1LL << lg;
      if ((r & (G - 1)) == 0)
      {
        const long long e = min((long long)rows, r + G);
        loc[lg] = max(loc[lg], (rp[e] - start) + (start & 3u));
      }
    }
  }
#pragma unroll
  for (int lg = 0; lg < 9; ++lg) if (loc[lg]) atomicMax(&s_max[lg], loc[lg]);
  __syncthreads();
  if (threadIdx.x < 9 && s_max[threadIdx.x]) atomicMax(&max_fill[threadIdx.x], s_max[threadIdx.x]);
}

__global__ void uniform_blocks_kernel(int rows, int G, int nb, u32 *blk)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i <= nb; i += (long long)gridDim.x * blockDim.x)
    blk[i] = (u32)min((long long)rows, i * G);
}

// returns the uniform group size (0: none usable)
static ViennaCLStatus uniform_group_size(ViennaCLBackend b, int rows, const u32 *row_ptr, int *G_out)
{
  *G_out = 0;
  if (getenv("VCL_B200_HOST_ROW_BLOCKS")) return ViennaCLSuccess;
  unsigned int *d_max = reinterpret_cast<unsigned int*>(b->dscal + 44);
  VCL_CUDA(b, cudaMemsetAsync(d_max, 0, 9 * sizeof(unsigned int), b->stream));
  row_group_fill_kernel<<<std::max(1, std::min(vcl_div_up(rows, 256), b->sm_count * 8)), 256, 0, b->stream>>>(rows, row_ptr, d_max);
  VCL_LAUNCHED(b, "row_group_fill_kernel");
  unsigned int h[9];
  VCL_CUDA(b, cudaMemcpyAsync(h, d_max, sizeof(h), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  static_assert(VCL_B200_CSR_BLOCK_ROWS <= 256 && (VCL_B200_CSR_BLOCK_ROWS & (VCL_B200_CSR_BLOCK_ROWS - 1)) == 0, "row blocks: a power of two <= 256 rows");
  int top = 0;
  while ((1 << top) < VCL_B200_CSR_BLOCK_ROWS) ++top;
  for (int lg = top; lg >= 5; --lg)
    if (h[lg] <= (unsigned int)VCL_B200_CSR_BLOCK_NNZ) { *G_out = 1 << lg; break; }
  return ViennaCLSuccess;
}

extern "C" ViennaCLStatus ViennaCLCUDAcsr_row_blocks(ViennaCLBackend b, ViennaCLInt rows, const unsigned int *row_ptr,
                                                     unsigned int *row_blocks, ViennaCLInt *num_blocks)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, rows >= 0 && num_blocks, "bad arguments");
  if (rows == 0) { *num_blocks = 0; return ViennaCLSuccess; }
  VCL_REQUIRE(b, row_ptr, "null row_ptr");
  int G = 0;
  VCL_TRY(uniform_group_size(b, rows, row_ptr, &G));
  if (G > 0)
  {
    const int nb = vcl_div_up(rows, G);
    if (row_blocks)
    {
      VCL_REQUIRE(b, *num_blocks >= nb, "row_blocks buffer too small");
      uniform_blocks_kernel<<<std::max(1, std::min(vcl_div_up(nb + 1, 256), b->sm_count * 4)), 256, 0, b->stream>>>(rows, G, nb, row_blocks);
      VCL_LAUNCHED(b, "uniform_blocks_kernel");
      vcl_plan_register(b, row_ptr, rows, row_blocks, nb);
    }
    *num_blocks = nb;
    return ViennaCLSuccess;
  }
  std::vector<u32> rp((size_t)rows + 1), blk;
  VCL_CUDA(b, cudaMemcpyAsync(rp.data(), row_ptr, sizeof(u32) * ((size_t)rows + 1), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  build_row_blocks(rp.data(), rows, blk);
  const int nb = (int)blk.size() - 1;
  if (row_blocks)
  {
    VCL_REQUIRE(b, *num_blocks >= nb, "row_blocks buffer too small");
    VCL_CUDA(b, cudaMemcpyAsync(row_blocks, blk.data(), sizeof(u32) * blk.size(), cudaMemcpyHostToDevice, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    vcl_plan_register(b, row_ptr, rows, row_blocks, nb);
  }
  *num_blocks = nb;
  return ViennaCLSuccess;
}

#endif // !VCL_F32

// ------------------------------------------------------------------------------------------------
// prod_impl
// ------------------------------------------------------------------------------------------------
extern "C" ViennaCLStatus ViennaCLCUDADcsrmv(ViennaCLBackend b, ViennaCLInt rows, ViennaCLInt cols, ViennaCLInt nnz,
                                             const unsigned int *row_ptr, const unsigned int *col_idx, const real *values,
                                             const unsigned int *row_blocks, ViennaCLInt num_blocks,
                                             const real *x, ViennaCLInt offx, ViennaCLInt incx, real alpha,
                                             real *y, ViennaCLInt offy, ViennaCLInt incy, real beta)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, rows >= 0 && cols >= 0 && nnz >= 0, "negative size");
  if (rows == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, row_ptr && x && y && (nnz == 0 || (col_idx && values)), "null pointer");
  VCL_REQUIRE(b, incx != 0 && incy != 0, "zero stride");
  VCL_REQUIRE(b, x != y, "x and y alias: the facade resolves x = A*x through a temporary (compressed_matrix.hpp:1237-1242)");
  ViennaCLCUDADcsr A = {rows, cols, nnz, row_ptr, col_idx, values, row_blocks, num_blocks};
  EpiAxpby epi = {y, offy, incy, alpha, beta};
  XVec xv = make_xvec(x, offx, incx);
  return vcl_launch_csr(b, A, xv, epi);
}

extern "C" ViennaCLStatus ViennaCLCUDADsellmv(ViennaCLBackend b, ViennaCLInt rows, ViennaCLInt cols, ViennaCLInt rows_per_block,
                                              const unsigned int *columns_per_block, const unsigned int *col_idx,
                                              const unsigned int *block_start, const real *values,
                                              const real *x, ViennaCLInt offx, ViennaCLInt incx, real alpha,
                                              real *y, ViennaCLInt offy, ViennaCLInt incy, real beta)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, rows >= 0 && cols >= 0 && rows_per_block > 0, "bad size");
  if (rows == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, columns_per_block && block_start && x && y, "null pointer");
  VCL_REQUIRE(b, incx != 0 && incy != 0, "zero stride");
  VCL_REQUIRE(b, x != y, "x and y alias");
  ViennaCLCUDADsell A = {rows, cols, rows_per_block, columns_per_block, col_idx, block_start, values};
  EpiAxpby epi = {y, offy, incy, alpha, beta};
  epi.sell_f32 = sizeof(real) == 4;
  XVec xv = make_xvec(x, offx, incx);
  return vcl_launch_sell(b, A, xv, epi);
}

// ------------------------------------------------------------------------------------------------
// row_info (cuda/sparse_matrix_operations.hpp:53-119): inf-/1-/2-norm or diagonal of every row
// ------------------------------------------------------------------------------------------------
__global__ void csr_row_info_kernel(int rows, const u32 * __restrict__ rp, const u32 * __restrict__ ci,
                                    const real * __restrict__ va, real *out, int option)
{
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x)
  {
    real value = 0.0;
    const u32 e = rp[r + 1];
    switch (option)
    {
    case 0: for (u32 k = rp[r]; k < e; ++k) value = fmax(value, fabs(va[k])); break;
    case 1: for (u32 k = rp[r]; k < e; ++k) value += fabs(va[k]); break;
    case 2: for (u32 k = rp[r]; k < e; ++k) value += va[k] * va[k]; value = sqrt(value); break;
    default:
      for (u32 k = rp[r]; k < e; ++k)
        if (ci[k] == (u32)r) { value = va[k]; break; }
      break;
    }
    out[r] = value;
  }
}

extern "C" ViennaCLStatus ViennaCLCUDADcsr_row_info(ViennaCLBackend b, ViennaCLInt rows,
                                                    const unsigned int *row_ptr, const unsigned int *col_idx, const real *values,
                                                    real *result, ViennaCLInt option)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, rows >= 0 && option >= 0 && option <= 3, "bad arguments");
  if (rows == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, row_ptr && result, "null pointer");
  int grid = std::min(vcl_div_up(rows, 256), b->sm_count * 8);
  csr_row_info_kernel<<<grid, 256, 0, b->stream>>>(rows, row_ptr, col_idx, values, result, option);
  VCL_LAUNCHED(b, "csr_row_info_kernel");
  return ViennaCLSuccess;
}

// ------------------------------------------------------------------------------------------------
// CSR -> SELL-C (sliced_ell_matrix.hpp:140-214 layout): widths on device, slice offsets scanned on the host (set-up path),
// entries scattered on device.
// ------------------------------------------------------------------------------------------------
__global__ void sell_width_kernel(int rows, int C, const u32 * __restrict__ rp, u32 *cpb)
{
  const int nslices = (rows - 1) / C + 1;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < nslices; s += (long long)gridDim.x * blockDim.x)
  {
    u32 w = 0;
    const long long r_end = min((long long)rows, (s + 1) * C);
    for (long long r = s * C; r < r_end; ++r) w = max(w, rp[r + 1] - rp[r]);
    cpb[s] = w;
  }
}

__global__ void sell_fill_kernel(int rows, int C, const u32 * __restrict__ rp, const u32 * __restrict__ cci, const real * __restrict__ cva,
                                 const u32 * __restrict__ cpb, const u32 * __restrict__ bs, u32 *ci, real *va)
{
  // one thread per (row of the padded slice): writes its real entries, then zero/col-0 padding up to the slice width
  const long long padded_rows = ((long long)(rows - 1) / C + 1) * C;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < padded_rows; r += (long long)gridDim.x * blockDim.x)
  {
    const u32 s = (u32)(r / C);
    const u32 w = cpb[s];
    size_t idx = (size_t)bs[s] + (size_t)(r - (long long)s * C);
    u32 j = 0;
    if (r < rows)
    {
      const u32 e = rp[r + 1];
      for (u32 k = rp[r]; k < e; ++k, ++j, idx += C) { ci[idx] = cci[k]; va[idx] = cva[k]; }
    }
    for (; j < w; ++j, idx += C) { ci[idx] = 0u; va[idx] = 0.0; }
  }
}

extern "C" ViennaCLStatus ViennaCLCUDADcsr2sell(ViennaCLBackend b, ViennaCLInt rows, ViennaCLInt C,
                                                const unsigned int *row_ptr, const unsigned int *csr_col, const real *csr_val,
                                                unsigned int *columns_per_block, unsigned int *block_start, long long *padded_nnz,
                                                unsigned int *col_idx, real *values)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, rows >= 0 && C > 0 && padded_nnz, "bad arguments");
  if (rows == 0) { *padded_nnz = 0; return ViennaCLSuccess; }
  VCL_REQUIRE(b, row_ptr && columns_per_block && block_start, "null pointer");
  const int nslices = (rows - 1) / C + 1;
  if (!col_idx || !values)
  {
    int grid = std::min(vcl_div_up(nslices, 256), b->sm_count * 8);
    sell_width_kernel<<<grid, 256, 0, b->stream>>>(rows, C, row_ptr, columns_per_block);
    VCL_LAUNCHED(b, "sell_width_kernel");
    std::vector<u32> w((size_t)nslices), start((size_t)nslices);
    VCL_CUDA(b, cudaMemcpyAsync(w.data(), columns_per_block, sizeof(u32) * nslices, cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    unsigned long long off = 0;
    for (int s = 0; s < nslices; ++s) { start[s] = (u32)off; off += (unsigned long long)w[s] * (unsigned long long)C; }
    VCL_REQUIRE(b, off <= 0xFFFFFFFFull, "SELL storage exceeds 32-bit offsets (sliced_ell_matrix.hpp:134-137 uses unsigned int)");
    VCL_CUDA(b, cudaMemcpyAsync(block_start, start.data(), sizeof(u32) * nslices, cudaMemcpyHostToDevice, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    *padded_nnz = (long long)off;
    return ViennaCLSuccess;
  }
  const long long padded_rows = (long long)nslices * C;
  int grid = std::min(vcl_div_up(padded_rows, 256), b->sm_count * 8);
  sell_fill_kernel<<<grid, 256, 0, b->stream>>>(rows, C, row_ptr, csr_col, csr_val, columns_per_block, block_start, col_idx, values);
  VCL_LAUNCHED(b, "sell_fill_kernel");
  return ViennaCLSuccess;
}

// ------------------------------------------------------------------------------------------------
// CSR -> SELL-C-sigma.  One CTA per window of sigma rows: (length, local index) keys in shared memory, bitonic sort by
// decreasing length with the row index as tie-break (= a stable sort), then the sorted window is written to row_perm.
// Widths and entries are then taken through the permutation; the layout inside a slice is that of sigma = 1.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sell_sigma_sort_kernel(int rows, int padded_rows, int sigma, int pow2, const u32 * __restrict__ rp, u32 *perm)
{
  extern __shared__ unsigned long long s_key[];            // pow2 keys: (0xFFFFFFFF - length) << 32 | local index; padding sorts last
  const long long w0 = (long long)blockIdx.x * sigma;
  for (int i = threadIdx.x; i < pow2; i += blockDim.x)
  {
    const long long r = w0 + i;
    unsigned long long key = ~0ull;
    if (i < sigma && r < rows) key = ((unsigned long long)(0xFFFFFFFFu - (rp[r + 1] - rp[r])) << 32) | (unsigned)i;
    s_key[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= pow2; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1)
    {
      for (int i = threadIdx.x; i < pow2; i += blockDim.x)
      {
        const int j = i ^ stride;
        if (j > i)
        {
          const bool up = (i & size) == 0;
          const unsigned long long a = s_key[i], c = s_key[j];
          if ((a > c) == up) { s_key[i] = c; s_key[j] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < sigma; i += blockDim.x)
  {
    const long long pos = w0 + i;
    if (pos >= padded_rows) break;
    const unsigned long long key = s_key[i];
    perm[pos] = (key == ~0ull) ? 0xFFFFFFFFu : (u32)(w0 + (long long)(key & 0xFFFFFFFFull));
  }
}

__global__ void sell_width_perm_kernel(int rows, int C, const u32 * __restrict__ rp, const u32 * __restrict__ perm, u32 *cpb)
{
  const int nslices = (rows - 1) / C + 1;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < nslices; s += (long long)gridDim.x * blockDim.x)
  {
    u32 w = 0;
    for (long long i = s * C; i < (s + 1) * C; ++i)
    {
      const u32 r = perm[i];
      if (r != 0xFFFFFFFFu) w = max(w, rp[r + 1] - rp[r]);
    }
    cpb[s] = w;
  }
}

__global__ void sell_fill_perm_kernel(int rows, int C, const u32 * __restrict__ rp, const u32 * __restrict__ cci, const real * __restrict__ cva,
                                      const u32 * __restrict__ perm, const u32 * __restrict__ cpb, const u32 * __restrict__ bs, u32 *ci, real *va)
{
  const long long padded_rows = ((long long)(rows - 1) / C + 1) * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < padded_rows; i += (long long)gridDim.x * blockDim.x)
  {
    const u32 s = (u32)(i / C);
    const u32 w = cpb[s];
    size_t idx = (size_t)bs[s] + (size_t)(i - (long long)s * C);
    const u32 r = perm[i];
    u32 j = 0;
    if (r != 0xFFFFFFFFu)
    {
      const u32 e = rp[r + 1];
      for (u32 k = rp[r]; k < e; ++k, ++j, idx += C) { ci[idx] = cci[k]; va[idx] = cva[k]; }
    }
    for (; j < w; ++j, idx += C) { ci[idx] = 0u; va[idx] = 0.0; }
  }
}

extern "C" ViennaCLStatus ViennaCLCUDADcsr2sell_sigma(ViennaCLBackend b, ViennaCLInt rows, ViennaCLInt C, ViennaCLInt sigma,
                                                      const unsigned int *row_ptr, const unsigned int *csr_col, const real *csr_val,
                                                      unsigned int *row_perm, unsigned int *columns_per_block, unsigned int *block_start,
                                                      long long *padded_nnz, unsigned int *col_idx, real *values)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, rows >= 0 && C > 0 && padded_nnz, "bad arguments");
  VCL_REQUIRE(b, sigma >= C && sigma % C == 0 && sigma <= 4096, "sigma must be a multiple of rows_per_block and at most 4096");
  if (rows == 0) { *padded_nnz = 0; return ViennaCLSuccess; }
  VCL_REQUIRE(b, row_ptr && row_perm && columns_per_block && block_start, "null pointer");
  const int nslices = (rows - 1) / C + 1;
  const int padded_rows = nslices * C;
  if (!col_idx || !values)
  {
    int pow2 = 1;
    while (pow2 < sigma) pow2 <<= 1;
    const int nwin = vcl_div_up(padded_rows, sigma);
    sell_sigma_sort_kernel<<<nwin, 256, (size_t)pow2 * sizeof(unsigned long long), b->stream>>>(rows, padded_rows, sigma, pow2, row_ptr, row_perm);
    VCL_LAUNCHED(b, "sell_sigma_sort_kernel");
    int grid = std::min(vcl_div_up(nslices, 256), b->sm_count * 8);
    sell_width_perm_kernel<<<grid, 256, 0, b->stream>>>(rows, C, row_ptr, row_perm, columns_per_block);
    VCL_LAUNCHED(b, "sell_width_perm_kernel");
    std::vector<u32> w((size_t)nslices), start((size_t)nslices);
    VCL_CUDA(b, cudaMemcpyAsync(w.data(), columns_per_block, sizeof(u32) * nslices, cudaMemcpyDeviceToHost, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    unsigned long long off = 0;
    for (int s = 0; s < nslices; ++s) { start[s] = (u32)off; off += (unsigned long long)w[s] * (unsigned long long)C; }
    VCL_REQUIRE(b, off <= 0xFFFFFFFFull, "SELL storage exceeds 32-bit offsets (sliced_ell_matrix.hpp:134-137 uses unsigned int)");
    VCL_CUDA(b, cudaMemcpyAsync(block_start, start.data(), sizeof(u32) * nslices, cudaMemcpyHostToDevice, b->stream));
    VCL_CUDA(b, cudaStreamSynchronize(b->stream));
    *padded_nnz = (long long)off;
    return ViennaCLSuccess;
  }
  int grid = std::min(vcl_div_up(padded_rows, 256), b->sm_count * 8);
  sell_fill_perm_kernel<<<grid, 256, 0, b->stream>>>(rows, C, row_ptr, csr_col, csr_val, row_perm, columns_per_block, block_start, col_idx, values);
  VCL_LAUNCHED(b, "sell_fill_perm_kernel");
  return ViennaCLSuccess;
}

extern "C" ViennaCLStatus ViennaCLCUDADsellmv_struct(ViennaCLBackend b, const ViennaCLCUDADsell *A,
                                                     const real *x, ViennaCLInt offx, ViennaCLInt incx, real alpha,
                                                     real *y, ViennaCLInt offy, ViennaCLInt incy, real beta)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, A != nullptr && A->rows >= 0 && A->cols >= 0 && A->rows_per_block > 0, "bad matrix");
  if (A->rows == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, A->columns_per_block && A->block_start && x && y, "null pointer");
  VCL_REQUIRE(b, incx != 0 && incy != 0, "zero stride");
  VCL_REQUIRE(b, x != y, "x and y alias");
  EpiAxpby epi = {y, offy, incy, alpha, beta};
  epi.sell_f32 = sizeof(real) == 4;
  XVec xv = make_xvec(x, offx, incx);
  return vcl_launch_sell(b, *A, xv, epi);
}

// ------------------------------------------------------------------------------------------------
// CSR -> ELL (ell_matrix.hpp:122-166) and CSR -> HYB (hyb_matrix.hpp:127-214), AlignmentV = 1.  Row lengths are analysed
// on the host from a D2H copy of row_ptr (set-up path, like the reference, which builds both formats entirely on the
// host); the entries are scattered on the device.
// ------------------------------------------------------------------------------------------------
__global__ void ell_fill_kernel(int rows, int width, const u32 * __restrict__ rp, const u32 * __restrict__ cci, const real * __restrict__ cva,
                                u32 *coords, real *elements, const u32 * __restrict__ tail_rows, u32 *tail_cols, real *tail_elements)
{
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x)
  {
    const u32 b0 = rp[r], e = rp[r + 1];
    size_t idx = (size_t)r;
    u32 k = b0;
    int j = 0;
    for (; j < width && k < e; ++j, ++k, idx += (size_t)rows) { coords[idx] = cci[k]; elements[idx] = cva[k]; }
    for (; j < width; ++j, idx += (size_t)rows) { coords[idx] = 0u; elements[idx] = 0.0; }
    if (tail_rows)
    {
      u32 t = tail_rows[r];
      for (; k < e; ++k, ++t) { tail_cols[t] = cci[k]; tail_elements[t] = cva[k]; }
    }
  }
}

static ViennaCLStatus fetch_row_ptr(ViennaCLBackend b, int rows, const u32 *row_ptr, std::vector<u32> &rp)
{
  rp.resize((size_t)rows + 1);
  VCL_CUDA(b, cudaMemcpyAsync(rp.data(), row_ptr, sizeof(u32) * ((size_t)rows + 1), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  return ViennaCLSuccess;
}

extern "C" ViennaCLStatus ViennaCLCUDADcsr2ell(ViennaCLBackend b, ViennaCLInt rows, const unsigned int *row_ptr,
                                               const unsigned int *csr_col, const real *csr_val, ViennaCLInt *maxnnz,
                                               unsigned int *coords, real *elements)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, rows >= 0 && maxnnz, "bad arguments");
  if (rows == 0) { *maxnnz = 0; return ViennaCLSuccess; }
  VCL_REQUIRE(b, row_ptr, "null row_ptr");
  if (!coords || !elements)
  {
    std::vector<u32> rp;
    VCL_TRY(fetch_row_ptr(b, rows, row_ptr, rp));
    u32 w = 0;
    for (int r = 0; r < rows; ++r) w = std::max(w, rp[r + 1] - rp[r]);
    VCL_REQUIRE(b, (unsigned long long)w * (unsigned long long)rows <= 0x7FFFFFFFull, "ELL storage exceeds 32-bit sizes");
    *maxnnz = (ViennaCLInt)w;
    return ViennaCLSuccess;
  }
  if (*maxnnz == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, csr_col && csr_val, "null CSR array");
  ell_fill_kernel<<<std::min(vcl_div_up(rows, 256), b->sm_count * 8), 256, 0, b->stream>>>(rows, *maxnnz, row_ptr, csr_col, csr_val, coords, elements,
                                                                                         nullptr, nullptr, nullptr);
  VCL_LAUNCHED(b, "ell_fill_kernel");
  return ViennaCLSuccess;
}

extern "C" ViennaCLStatus ViennaCLCUDADcsr2hyb(ViennaCLBackend b, ViennaCLInt rows, ViennaCLInt cols, const unsigned int *row_ptr,
                                               const unsigned int *csr_col, const real *csr_val, real csr_threshold,
                                               ViennaCLInt *ell_width, ViennaCLInt *csr_nnz,
                                               unsigned int *ell_coords, real *ell_elements,
                                               unsigned int *csr_rows, unsigned int *csr_cols, real *csr_elements)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, rows >= 0 && cols >= 0 && ell_width && csr_nnz, "bad arguments");
  if (rows == 0) { *ell_width = 0; *csr_nnz = 0; return ViennaCLSuccess; }
  VCL_REQUIRE(b, row_ptr, "null row_ptr");
  std::vector<u32> rp;
  VCL_TRY(fetch_row_ptr(b, rows, row_ptr, rp));
  if (!ell_coords || !csr_rows)
  {
    // hyb_matrix.hpp:139-166: histogram of row lengths, smallest width covering >= threshold of the rows
    u32 maxw = 0;
    for (int r = 0; r < rows; ++r) maxw = std::max(maxw, rp[r + 1] - rp[r]);
    std::vector<size_t> hist((size_t)maxw + 1, 0);
    for (int r = 0; r < rows; ++r) hist[rp[r + 1] - rp[r]] += 1;
    size_t sum = 0; u32 w = maxw;
    for (u32 ind = 0; ind <= maxw; ++ind)
    {
      sum += hist[ind];
      if ((real)sum >= csr_threshold * (real)rows) { w = ind; break; }
    }
    unsigned long long tail = 0;
    for (int r = 0; r < rows; ++r) if (rp[r + 1] - rp[r] > w) tail += (rp[r + 1] - rp[r]) - w;
    VCL_REQUIRE(b, (unsigned long long)w * (unsigned long long)rows <= 0x7FFFFFFFull && tail <= 0x7FFFFFFFull, "HYB storage exceeds 32-bit sizes");
    *ell_width = (ViennaCLInt)w;
    *csr_nnz = (ViennaCLInt)(tail > 0 ? tail : 1);            // one dummy entry when empty (hyb_matrix.hpp:209-213)
    return ViennaCLSuccess;
  }
  VCL_REQUIRE(b, csr_cols && csr_elements && (*ell_width == 0 || ell_elements), "null output array");
  const u32 w = (u32)*ell_width;
  std::vector<u32> tr((size_t)rows + 1);
  u32 t = 0;
  for (int r = 0; r < rows; ++r) { tr[r] = t; if (rp[r + 1] - rp[r] > w) t += (rp[r + 1] - rp[r]) - w; }
  tr[rows] = t;
  VCL_CUDA(b, cudaMemcpyAsync(csr_rows, tr.data(), sizeof(u32) * ((size_t)rows + 1), cudaMemcpyHostToDevice, b->stream));
  if (t == 0)
  {
    VCL_CUDA(b, cudaMemsetAsync(csr_cols, 0, sizeof(u32), b->stream));
    VCL_CUDA(b, cudaMemsetAsync(csr_elements, 0, sizeof(real), b->stream));
  }
  VCL_REQUIRE(b, rp[rows] == 0 || (csr_col && csr_val), "null CSR array");
  ell_fill_kernel<<<std::min(vcl_div_up(rows, 256), b->sm_count * 8), 256, 0, b->stream>>>(rows, (int)w, row_ptr, csr_col, csr_val, ell_coords, ell_elements,
                                                                                         csr_rows, csr_cols, csr_elements);
  VCL_LAUNCHED(b, "ell_fill_kernel");
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));              // tr (pageable) must outlive the copy
  return ViennaCLSuccess;
}

#ifndef VCL_F32      // index-only helper: compiled once (double build)
// ------------------------------------------------------------------------------------------------
// COO -> CSR index (coordinate_matrix.hpp:47-102 stores (row, col) pairs sorted by row).  The reference's CUDA kernel
// (cuda/sparse_matrix_operations.hpp:1239-1340) runs a segmented reduction over 64 fixed groups; here the entries get a
// row pointer once, and every product / solver step streams them through the CSR kernels (12 instead of 16 bytes per entry).
// ------------------------------------------------------------------------------------------------
__global__ void coo_index_kernel(int rows, int nnz, const u32 * __restrict__ coords, u32 *row_ptr, u32 *col_idx, int *unsorted)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i <= nnz; i += (long long)gridDim.x * blockDim.x)
  {
    const long long prev = i > 0 ? (long long)coords[2 * (i - 1)] : -1;
    const long long cur = i < nnz ? (long long)coords[2 * i] : (long long)rows;
    if (cur < prev || (i < nnz && cur >= rows)) { *unsorted = 1; continue; }
    for (long long q = prev + 1; q <= cur; ++q) row_ptr[q] = (u32)i;      // rows (prev, cur] start at entry i
    if (i < nnz) col_idx[i] = coords[2 * i + 1];
  }
}

extern "C" ViennaCLStatus ViennaCLCUDAcoo2csr(ViennaCLBackend b, ViennaCLInt rows, ViennaCLInt nnz, const unsigned int *coords,
                                              unsigned int *row_ptr, unsigned int *col_idx)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, rows >= 0 && nnz >= 0 && row_ptr && (nnz == 0 || (coords && col_idx)), "bad arguments");
  int *flag = reinterpret_cast<int*>(b->dscal + 40);
  VCL_CUDA(b, cudaMemsetAsync(flag, 0, sizeof(int), b->stream));
  coo_index_kernel<<<std::max(1, std::min(vcl_div_up((long long)nnz + 1, 256), b->sm_count * 8)), 256, 0, b->stream>>>(rows, nnz, coords, row_ptr, col_idx, flag);
  VCL_LAUNCHED(b, "coo_index_kernel");
  int h = 0;
  VCL_CUDA(b, cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  VCL_CUDA(b, cudaStreamSynchronize(b->stream));
  VCL_REQUIRE(b, h == 0, "coordinate_matrix entries must be sorted by row with row indices < size1 (coordinate_matrix.hpp:72-88 produces them so)");
  return ViennaCLSuccess;
}

#endif // !VCL_F32

extern "C" ViennaCLStatus ViennaCLCUDADcoomv(ViennaCLBackend b, ViennaCLInt rows, ViennaCLInt cols, ViennaCLInt nnz,
                                             const unsigned int *row_ptr, const unsigned int *col_idx, const real *elements,
                                             const unsigned int *row_blocks, ViennaCLInt num_blocks,
                                             const real *x, ViennaCLInt offx, ViennaCLInt incx, real alpha,
                                             real *y, ViennaCLInt offy, ViennaCLInt incy, real beta)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, rows >= 0 && cols >= 0 && nnz >= 0, "negative size");
  if (rows == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, row_ptr && x && y && (nnz == 0 || (col_idx && elements)), "null pointer");
  VCL_REQUIRE(b, incx != 0 && incy != 0 && x != y, "bad vectors");
  ViennaCLCUDADcsr A = {rows, cols, nnz, row_ptr, col_idx, elements, row_blocks, num_blocks};
  EpiCoo epi = {y, offy, incy, alpha, beta};
  return vcl_launch_csr(b, A, make_xvec(x, offx, incx), epi);
}

} // namespace VCL_NS
