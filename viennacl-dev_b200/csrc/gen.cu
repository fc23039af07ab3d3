// gen.cu -- device-side generators for the synthetic benchmark inputs.
// Generalises viennacl/tools/matrix_generation.hpp:47-88 (2-D 5-point Laplacian assembled through a std::map on the host,
// 2.6 s per million rows) to 2-D/3-D 5-/7-point stencils with first-order upwind convection, written straight into CSR
// arrays in device memory: row offsets have a closed form, so every row is emitted independently.
// Conventions (identical to oracle/vcl_oracle.c vclo_gen_stencil*): row = i + nx*(j + ny*l); neighbours outside the grid
// are dropped; columns ascend within a row; down/south/west = -1-c, diagonal = 2*dim + sum(c), east/north/up = -1.
#include "common.cuh"
#include "prec.cuh"
#include <algorithm>

namespace VCL_NS
{

struct StencilGeom { int nx, ny, nz; real cx, cy, cz; };

// number of stored entries in rows [0, r)
__device__ __forceinline__ unsigned long long stencil_offset(const StencilGeom &g, long long r)
{
  const long long nxy = (long long)g.nx * g.ny;
  const long long l = r / nxy, rem = r - l * nxy, j = rem / g.nx, i = rem - j * g.nx;
  const long long L = j + (long long)g.ny * l;                               // complete x-lines before r
  long long missing = (L + (i > 0 ? 1 : 0))                                  // rows with i' == 0
                    + L                                                      // rows with i' == nx-1
                    + (l * g.nx + (j > 0 ? g.nx : i))                        // rows with j' == 0
                    + (l * g.nx + (j == g.ny - 1 ? i : 0));                  // rows with j' == ny-1
  long long per_row = 5;
  if (g.nz > 1)
  {
    per_row = 7;
    missing += (l > 0 ? nxy : rem);                                          // rows with l' == 0
    missing += (l > g.nz - 1 ? nxy : (l == g.nz - 1 ? rem : 0));             // rows with l' == nz-1 (l == nz only for r == rows)
  }
  return (unsigned long long)(per_row * r - missing);
}

__global__ void __launch_bounds__(256)
stencil_kernel(StencilGeom g, long long row_begin, long long row_end, u32 *rp, u32 *ci, real *va)
{
  const long long nxy = (long long)g.nx * g.ny;
  const unsigned long long base = stencil_offset(g, row_begin);
  for (long long r = row_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= row_end; r += (long long)gridDim.x * blockDim.x)
  {
    unsigned long long k = stencil_offset(g, r) - base;
    rp[r - row_begin] = (u32)k;
    if (r == row_end) break;
    const long long l = r / nxy, rem = r - l * nxy, j = rem / g.nx, i = rem - j * g.nx;
    const bool three_d = g.nz > 1;
    if (three_d && l > 0)   { ci[k] = (u32)(r - nxy);  va[k] = -1.0 - g.cz; ++k; }
    if (j > 0)              { ci[k] = (u32)(r - g.nx); va[k] = -1.0 - g.cy; ++k; }
    if (i > 0)              { ci[k] = (u32)(r - 1);    va[k] = -1.0 - g.cx; ++k; }
    ci[k] = (u32)r; va[k] = (three_d ? 6.0 + g.cx + g.cy + g.cz : 4.0 + g.cx + g.cy); ++k;
    if (i < g.nx - 1)       { ci[k] = (u32)(r + 1);    va[k] = -1.0; ++k; }
    if (j < g.ny - 1)       { ci[k] = (u32)(r + g.nx); va[k] = -1.0; ++k; }
    if (three_d && l < g.nz - 1) { ci[k] = (u32)(r + nxy); va[k] = -1.0; ++k; }
  }
}

static unsigned long long host_offset(const StencilGeom &g, long long r)
{
  const long long nxy = (long long)g.nx * g.ny;
  const long long l = r / nxy, rem = r - l * nxy, j = rem / g.nx, i = rem - j * g.nx;
  const long long L = j + (long long)g.ny * l;
  long long missing = (L + (i > 0 ? 1 : 0)) + L + (l * g.nx + (j > 0 ? g.nx : i)) + (l * g.nx + (j == g.ny - 1 ? i : 0));
  long long per_row = 5;
  if (g.nz > 1) { per_row = 7; missing += (l > 0 ? nxy : rem); missing += (l > g.nz - 1 ? nxy : (l == g.nz - 1 ? rem : 0)); }
  return (unsigned long long)(per_row * r - missing);
}

__global__ void __launch_bounds__(256)
fill_uniform_kernel(long long n, real *x, unsigned long long seed, long long index_offset, real lo, real hi)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
  {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(i + index_offset + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    const real u = (real)(z >> 11) * (1.0 / 9007199254740992.0);
    x[i] = lo + (hi - lo) * u;
  }
}

extern "C" {

ViennaCLStatus ViennaCLCUDADgenerate_stencil_rows(ViennaCLBackend b, ViennaCLInt nx, ViennaCLInt ny, ViennaCLInt nz,
                                                  real cx, real cy, real cz, long long row_begin, long long row_end,
                                                  unsigned int *row_ptr, unsigned int *col_idx, real *values, long long *nnz)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, nx > 0 && ny > 0 && nz > 0, "grid dimensions must be positive");
  const long long rows = (long long)nx * ny * nz;
  VCL_REQUIRE(b, row_begin >= 0 && row_begin <= row_end && row_end <= rows, "bad row range");
  VCL_REQUIRE(b, rows < 0xFFFFFFFFLL, "column indices are 32-bit (compressed_matrix.hpp:1190-1197)");
  StencilGeom g = {nx, ny, nz, cx, cy, cz};
  const unsigned long long cnt = host_offset(g, row_end) - host_offset(g, row_begin);
  VCL_REQUIRE(b, cnt <= 0xFFFFFFFFULL, "row offsets are 32-bit");
  if (nnz) *nnz = (long long)cnt;
  if (!row_ptr || !col_idx || !values) return ViennaCLSuccess;
  const long long work = row_end - row_begin + 1;
  int grid = (int)std::min((work + 255) / 256, (long long)b->sm_count * 16);
  stencil_kernel<<<grid, 256, 0, b->stream>>>(g, row_begin, row_end, row_ptr, col_idx, values);
  VCL_LAUNCHED(b, "stencil_kernel");
  return ViennaCLSuccess;
}

ViennaCLStatus ViennaCLCUDADgenerate_stencil(ViennaCLBackend b, ViennaCLInt nx, ViennaCLInt ny, ViennaCLInt nz,
                                             real cx, real cy, real cz,
                                             unsigned int *row_ptr, unsigned int *col_idx, real *values,
                                             long long *rows, long long *nnz)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, nx > 0 && ny > 0 && nz > 0, "grid dimensions must be positive");
  const long long n = (long long)nx * ny * nz;
  if (rows) *rows = n;
  return ViennaCLCUDADgenerate_stencil_rows(b, nx, ny, nz, cx, cy, cz, 0, n, row_ptr, col_idx, values, nnz);
}

ViennaCLStatus ViennaCLCUDADfill_uniform(ViennaCLBackend b, long long n, real *x, unsigned long long seed,
                                         long long index_offset, real lo, real hi)
{
  VCL_CHECK_BACKEND(b);
  VCL_REQUIRE(b, n >= 0, "negative size");
  if (n == 0) return ViennaCLSuccess;
  VCL_REQUIRE(b, x, "null pointer");
  int grid = (int)std::min((n + 255) / 256, (long long)b->sm_count * 16);
  fill_uniform_kernel<<<grid, 256, 0, b->stream>>>(n, x, seed, index_offset, lo, hi);
  VCL_LAUNCHED(b, "fill_uniform_kernel");
  return ViennaCLSuccess;
}

} // extern "C"
} // namespace VCL_NS
