/* oracle/vcl_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, CPU restatement of the reference's algorithms on the SpMV + Krylov hot path.  It is the
 * checker the CUDA path is compared against; it is never linked into, called from, or shipped with
 * the product library.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * PINNING: every function below is checked in tests/test_oracle.py against the UNMODIFIED reference
 * compiled from /root/reference (oracle/_ref/libvcl_ref.so, built by oracle/Makefile from ref_shim.cpp)
 * and against the golden fixtures in tests/golden/ that were produced by that reference
 * (tests/golden/make_golden.py).  GMRES: the reference's *host* pipelined GMRES is defective
 * (host_based/iterative_operations.hpp:820-822 drops trailing entries); vclo_gmres restates the
 * intended semantics (those of the reference's CUDA kernels, cuda/iterative_operations.hpp:1597-1894) and is
 * pinned against (i) the reference's Householder GMRES (gmres.hpp:449-631) and (ii) the reference pipelined
 * host path with the documented one-line fix (oracle/_ref/libvcl_ref_gmresfix.so).
 *
 * ARITHMETIC: this file is compiled with -ffp-contract=off; wherever the reference, built by oracle/Makefile's recipe
 * (g++ -O3 -march=x86-64-v3, generic tuning), fuses a multiply-add, fma() is written out explicitly, and where it does not
 * (the in-row CSR chain `dot += a[i]*x[col[i]]`: GCC avoids FMA in reduction chains under generic tuning) it is not.
 * tests/test_oracle.py checks the SpMV forms BIT-FOR-BIT against that reference build; the CUDA kernels use the same
 * explicit operations (__dmul_rn/__dadd_rn vs fma).
 *
 * All citations are relative to /root/reference.
 */
#include "vcl_oracle.h"
#include <tgmath.h>   /* sqrt / fabs / fma resolve to the float versions in the -DVCLO_F32 build */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef vclo_u32 u32;

int vclo_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void vclo_set_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Generators.  Same conventions as viennacl/tools/matrix_generation.hpp:47-88 (row = i + j*nx, diagonal 4,
 * neighbours -1, out-of-grid neighbours dropped); columns are emitted in ascending order, which is the order
 * a std::map-backed host matrix hands them to copy() (compressed_matrix.hpp:49-105).
 * Upwind convection: west/south/down = -1-c, diagonal += c, east/north/up = -1.
 * Pass rp == NULL to only count the non-zeros.
 * ---------------------------------------------------------------------------------------------- */
long long vclo_gen_stencil2d(int nx, int ny, vreal cx, vreal cy, u32 *rp, u32 *ci, vreal *v)
{
  long long k = 0;
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i)
    {
      long long row = (long long)i + (long long)j * nx;
      if (rp) rp[row] = (u32)k;
      if (j > 0)      { if (rp) { ci[k] = (u32)(row - nx); v[k] = -1.0 - cy; } ++k; }
      if (i > 0)      { if (rp) { ci[k] = (u32)(row - 1);  v[k] = -1.0 - cx; } ++k; }
      if (rp) { ci[k] = (u32)row; v[k] = 4.0 + cx + cy; } ++k;
      if (i < nx - 1) { if (rp) { ci[k] = (u32)(row + 1);  v[k] = -1.0; } ++k; }
      if (j < ny - 1) { if (rp) { ci[k] = (u32)(row + nx); v[k] = -1.0; } ++k; }
    }
  if (rp) rp[(long long)nx * ny] = (u32)k;
  return k;
}

long long vclo_gen_stencil3d(int nx, int ny, int nz, vreal cx, vreal cy, vreal cz, u32 *rp, u32 *ci, vreal *v)
{
  long long k = 0;
  long long nxy = (long long)nx * ny;
  for (int l = 0; l < nz; ++l)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i)
      {
        long long row = (long long)i + (long long)j * nx + (long long)l * nxy;
        if (rp) rp[row] = (u32)k;
        if (l > 0)      { if (rp) { ci[k] = (u32)(row - nxy); v[k] = -1.0 - cz; } ++k; }
        if (j > 0)      { if (rp) { ci[k] = (u32)(row - nx);  v[k] = -1.0 - cy; } ++k; }
        if (i > 0)      { if (rp) { ci[k] = (u32)(row - 1);   v[k] = -1.0 - cx; } ++k; }
        if (rp) { ci[k] = (u32)row; v[k] = 6.0 + cx + cy + cz; } ++k;
        if (i < nx - 1) { if (rp) { ci[k] = (u32)(row + 1);   v[k] = -1.0; } ++k; }
        if (j < ny - 1) { if (rp) { ci[k] = (u32)(row + nx);  v[k] = -1.0; } ++k; }
        if (l < nz - 1) { if (rp) { ci[k] = (u32)(row + nxy); v[k] = -1.0; } ++k; }
      }
  if (rp) rp[nxy * nz] = (u32)k;
  return k;
}

/* Counter-based uniform numbers (splitmix64 of seed + index): reproducible in C, numpy and CUDA alike. */
void vclo_fill_uniform(vreal *x, long long n, unsigned long long seed, vreal lo, vreal hi)
{
  for (long long i = 0; i < n; ++i)
  {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    vreal u = (vreal)(z >> 11) * (1.0 / 9007199254740992.0);
    x[i] = lo + (hi - lo) * u;
  }
}

/* ------------------------------------------------------------------------------------------------
 * CSR SpMV: viennacl/linalg/host_based/sparse_matrix_operations.hpp:110-186.
 * In-row accumulation is sequential in storage order; beta == 0 means "do not read y".
 * ---------------------------------------------------------------------------------------------- */
void vclo_csr_spmv(int rows, const u32 *rp, const u32 *ci, const vreal *v,
                   const vreal *x, int offx, int incx, vreal alpha,
                   vreal *y, int offy, int incy, vreal beta)
{
#ifdef _OPENMP
  #pragma omp parallel for
#endif
  for (long row = 0; row < (long)rows; ++row)
  {
    vreal dot = 0;
    size_t row_end = rp[row + 1];
    size_t i = rp[row];
#ifdef VCLO_F32
    /* float instantiation of the reference build (pinned empirically, tests/test_oracle.py): GCC vectorises the gather
     * loop 4 wide with an in-order reduction -- products rounded, then added -- and contracts the scalar remainder loop
     * (the last (row length mod 4) entries) to fmaf -- in the loop version for x stride 1 only; the strided version and the
     * double instantiation are neither vectorised nor contracted. */
    size_t vec_end = (incx == 1) ? i + ((row_end - i) & ~(size_t)3) : row_end;
    for (; i < vec_end; ++i)
      dot += v[i] * x[(size_t)ci[i] * (size_t)incx + (size_t)offx];
    for (; i < row_end; ++i)
      dot = fma(v[i], x[(size_t)ci[i] * (size_t)incx + (size_t)offx], dot);
#else
    for (; i < row_end; ++i)
      dot += v[i] * x[(size_t)ci[i] * (size_t)incx + (size_t)offx];
#endif
    size_t idx = (size_t)row * (size_t)incy + (size_t)offy;
    if (beta < 0 || beta > 0)
      y[idx] = fma(beta, y[idx], alpha * dot);   /* the reference build fuses this one */
    else
      y[idx] = alpha * dot;
  }
}

/* ------------------------------------------------------------------------------------------------
 * SELL-C-sigma (sigma = 1) builder: viennacl/sliced_ell_matrix.hpp:140-214.
 * slice b covers rows [b*C, min((b+1)*C, rows)); width = longest row in slice; entry j of row r sits at
 * block_start[b] + j*C + (r mod C); padding has value 0 and column 0.
 * ---------------------------------------------------------------------------------------------- */
long long vclo_sell_padded_nnz(int rows, const u32 *rp, int C)
{
  long long tot = 0;
  for (int b0 = 0; b0 < rows; b0 += C)
  {
    u32 w = 0;
    for (int r = b0; r < rows && r < b0 + C; ++r)
      if (rp[r + 1] - rp[r] > w) w = rp[r + 1] - rp[r];
    tot += (long long)w * C;
  }
  return tot;
}

void vclo_sell_build(int rows, const u32 *rp, const u32 *ci, const vreal *v, int C,
                     u32 *cols_per_block, u32 *block_start, u32 *col_idx, vreal *elements)
{
  long long tot = vclo_sell_padded_nnz(rows, rp, C);
  memset(col_idx, 0, sizeof(u32) * (size_t)tot);
  memset(elements, 0, sizeof(vreal) * (size_t)tot);
  size_t off = 0;
  int b = 0;
  for (int b0 = 0; b0 < rows; b0 += C, ++b)
  {
    u32 w = 0;
    for (int r = b0; r < rows && r < b0 + C; ++r)
      if (rp[r + 1] - rp[r] > w) w = rp[r + 1] - rp[r];
    cols_per_block[b] = w;
    block_start[b] = (u32)off;
    for (int r = b0; r < rows && r < b0 + C; ++r)
    {
      u32 j = 0;
      for (u32 k = rp[r]; k < rp[r + 1]; ++k, ++j)
      {
        size_t idx = off + (size_t)j * (size_t)C + (size_t)(r - b0);
        col_idx[idx] = ci[k];
        elements[idx] = v[k];
      }
    }
    off += (size_t)w * (size_t)C;
  }
}

/* SELL SpMV: host_based/sparse_matrix_operations.hpp:1796-1858 (zero-valued slots never touch x).
 * Loops over the ceil(rows/C) slices that exist (the reference loops one further when rows % C == 0: SURVEY 8c-2). */
void vclo_sell_spmv(int rows, int C, const u32 *cols_per_block, const u32 *block_start,
                    const u32 *col_idx, const vreal *elements,
                    const vreal *x, int offx, int incx, vreal alpha,
                    vreal *y, int offy, int incy, vreal beta)
{
  long nb = rows > 0 ? ((long)rows - 1) / C + 1 : 0;
#ifdef _OPENMP
  #pragma omp parallel for
#endif
  for (long b = 0; b < nb; ++b)
  {
    for (int rib = 0; rib < C; ++rib)
    {
      long row = b * C + rib;
      if (row >= rows) break;
      vreal acc = 0;
      for (u32 j = 0; j < cols_per_block[b]; ++j)
      {
        size_t idx = (size_t)block_start[b] + (size_t)j * (size_t)C + (size_t)rib;
        vreal val = elements[idx];
        if (val > 0 || val < 0) acc = fma(x[(size_t)col_idx[idx] * (size_t)incx + (size_t)offx], val, acc);   /* fused in the reference build */
      }
      size_t yi = (size_t)row * (size_t)incy + (size_t)offy;
#ifdef VCLO_F32
      if (beta < 0 || beta > 0) y[yi] = fma(alpha, acc, beta * y[yi]);   /* float instantiation of the reference build: the other product is fused */
#else
      if (beta < 0 || beta > 0) y[yi] = fma(beta, y[yi], alpha * acc);
#endif
      else                      y[yi] = alpha * acc;
    }
  }
}

/* ----------------------------------------------------------------------------------------------
 * ELL (ell_matrix.hpp:122-166) and HYB (hyb_matrix.hpp:127-214) layouts and products, AlignmentV = 1.
 * ELL: entry j of row r at j*rows + r, width = longest row, padding value 0 / column 0.
 * HYB: ELL part of width w = the smallest row length L such that at least csr_threshold (0.8) of the rows have
 *      <= L entries; the remaining entries of longer rows go to a CSR tail (one dummy entry when the tail is empty).
 * Products: host_based/sparse_matrix_operations.hpp:1503-1538 (ELL), :1873-1927 (HYB) -- serial loops, zero-valued ELL
 * slots never touch x.  ARITHMETIC (pinned against the live reference build, tests/test_oracle.py): the conditional ELL
 * update `sum += x*val` is contracted to an fma by GCC, the unconditional CSR-tail update of HYB is not.
 * ---------------------------------------------------------------------------------------------- */
#ifndef ELL_FUSED
#define ELL_FUSED 1
#endif
static inline vreal ell_madd(vreal x, vreal v, vreal acc)
{
#if ELL_FUSED
  return fma(x, v, acc);
#else
  return acc + x * v;       /* two roundings (file is compiled with -ffp-contract=off) */
#endif
}

int vclo_ell_width(int rows, const u32 *rp)
{
  u32 w = 0;
  for (int r = 0; r < rows; ++r) if (rp[r + 1] - rp[r] > w) w = rp[r + 1] - rp[r];
  return (int)w;
}

void vclo_ell_build(int rows, const u32 *rp, const u32 *ci, const vreal *v, int width, u32 *coords, vreal *elements)
{
  memset(coords, 0, sizeof(u32) * (size_t)rows * (size_t)width);
  memset(elements, 0, sizeof(vreal) * (size_t)rows * (size_t)width);
  for (int r = 0; r < rows; ++r)
  {
    u32 j = 0;
    for (u32 k = rp[r]; k < rp[r + 1] && j < (u32)width; ++k, ++j)
    {
      coords[(size_t)j * (size_t)rows + (size_t)r] = ci[k];
      elements[(size_t)j * (size_t)rows + (size_t)r] = v[k];
    }
  }
}

int vclo_hyb_width(int rows, int cols, const u32 *rp, double threshold)
{
  int maxw = vclo_ell_width(rows, rp);
  long *hist = (long*)calloc((size_t)cols + 2, sizeof(long));
  for (int r = 0; r < rows; ++r) hist[rp[r + 1] - rp[r]] += 1;
  long sum = 0;
  int w = maxw;
  for (int ind = 0; ind <= maxw; ++ind)
  {
    sum += hist[ind];
    if ((double)sum >= threshold * (double)rows) { w = ind; break; }
  }
  free(hist);
  return w;
}

/* csr_rows[rows+1]; csr_cols / csr_elements sized vclo_hyb_tail_nnz() (>= 1: dummy entry when empty) */
long long vclo_hyb_tail_nnz(int rows, const u32 *rp, int width)
{
  long long t = 0;
  for (int r = 0; r < rows; ++r) if (rp[r + 1] - rp[r] > (u32)width) t += (long long)(rp[r + 1] - rp[r]) - width;
  return t > 0 ? t : 1;
}

void vclo_hyb_build(int rows, const u32 *rp, const u32 *ci, const vreal *v, int width,
                    u32 *ell_coords, vreal *ell_elements, u32 *csr_rows, u32 *csr_cols, vreal *csr_elements)
{
  vclo_ell_build(rows, rp, ci, v, width, ell_coords, ell_elements);
  u32 t = 0;
  for (int r = 0; r < rows; ++r)
  {
    csr_rows[r] = t;
    for (u32 k = rp[r] + (u32)width; k < rp[r + 1]; ++k, ++t) { csr_cols[t] = ci[k]; csr_elements[t] = v[k]; }
  }
  csr_rows[rows] = t;
  if (t == 0) { csr_cols[0] = 0; csr_elements[0] = 0; }
}

void vclo_hyb_spmv(int rows, int width, const u32 *ell_coords, const vreal *ell_elements,
                   const u32 *csr_rows, const u32 *csr_cols, const vreal *csr_elements,
                   const vreal *x, int offx, int incx, vreal alpha, vreal *y, int offy, int incy, vreal beta)
{
  for (int r = 0; r < rows; ++r)
  {
    vreal sum = 0;
    for (int j = 0; j < width; ++j)
    {
      size_t idx = (size_t)j * (size_t)rows + (size_t)r;
      vreal val = ell_elements[idx];
      if (val > 0 || val < 0) sum = ell_madd(x[(size_t)ell_coords[idx] * (size_t)incx + (size_t)offx], val, sum);
    }
    if (csr_rows)
    {
      u32 k = csr_rows[r], k1 = csr_rows[r + 1];
#ifdef VCLO_F32
      /* float instantiation: vectorised 4 wide (rounded products, in-order adds) + contracted scalar remainder, like vclo_csr_spmv */
      u32 kv = (incx == 1) ? k + ((k1 - k) & ~3u) : k1;
      for (; k < kv; ++k)
        sum = sum + x[(size_t)csr_cols[k] * (size_t)incx + (size_t)offx] * csr_elements[k];
      for (; k < k1; ++k)
        sum = fma(x[(size_t)csr_cols[k] * (size_t)incx + (size_t)offx], csr_elements[k], sum);
#else
      for (; k < k1; ++k)
        sum = sum + x[(size_t)csr_cols[k] * (size_t)incx + (size_t)offx] * csr_elements[k];   /* tail: NOT contracted in the reference build */
#endif
    }
    size_t yi = (size_t)r * (size_t)incy + (size_t)offy;
    if (beta < 0 || beta > 0) y[yi] = fma(beta, y[yi], alpha * sum);
    else                      y[yi] = alpha * sum;
  }
}

void vclo_ell_spmv(int rows, int width, const u32 *coords, const vreal *elements,
                   const vreal *x, int offx, int incx, vreal alpha, vreal *y, int offy, int incy, vreal beta)
{
  vclo_hyb_spmv(rows, width, coords, elements, NULL, NULL, NULL, x, offx, incx, alpha, y, offy, incy, beta);
}

/* COO (coordinate_matrix.hpp:47-102: (row, col) pairs in row-major order) product, host_based/sparse_matrix_operations.hpp:
 * 1222-1247: y is first scaled by beta (or cleared), then every entry adds (alpha * a) * x[col] IN STORAGE ORDER. */
#ifndef COO_FUSED
#define COO_FUSED 1
#endif
void vclo_coo_spmv(int rows, long long nnz, const u32 *coords, const vreal *elements,
                   const vreal *x, vreal alpha, vreal *y, vreal beta)
{
  if (beta < 0 || beta > 0) for (int i = 0; i < rows; ++i) y[i] *= beta;
  else                      for (int i = 0; i < rows; ++i) y[i] = 0;
  for (long long i = 0; i < nnz; ++i)
  {
    vreal t = alpha * elements[i];
#if COO_FUSED
    y[coords[2 * i]] = fma(t, x[coords[2 * i + 1]], y[coords[2 * i]]);
#else
    y[coords[2 * i]] = y[coords[2 * i]] + t * x[coords[2 * i + 1]];
#endif
  }
}

/* detail::row_info(A, vec, SPARSE_ROW_DIAGONAL): host_based/sparse_matrix_operations.hpp:52-98 (0 if absent). */
void vclo_csr_diag(int rows, const u32 *rp, const u32 *ci, const vreal *v, vreal *diag)
{
  for (int r = 0; r < rows; ++r)
  {
    vreal val = 0;
    for (u32 k = rp[r]; k < rp[r + 1]; ++k)
      if (ci[k] == (u32)r) { val = v[k]; break; }
    diag[r] = val;
  }
}

/* norm_2 / inner_prod: host_based/vector_operations.hpp:557-576, 463-472 -- plain sums, no scaling. */
vreal vclo_norm2(const vreal *x, long long n)
{
  vreal s = 0;
#ifdef _OPENMP
  #pragma omp parallel for reduction(+: s) if (n > 5000)
#endif
  for (long long i = 0; i < n; ++i) s += x[i] * x[i];
  return sqrt(s);
}
vreal vclo_inner_prod(const vreal *x, const vreal *y, long long n)
{
  vreal s = 0;
#ifdef _OPENMP
  #pragma omp parallel for reduction(+: s) if (n > 5000)
#endif
  for (long long i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}

/* Fused SpMV + dots: host_based/iterative_operations.hpp:58-103 (pipelined_prod_impl, CSR). */
static void fused_prod(int rows, const u32 *rp, const u32 *ci, const vreal *v,
                       const vreal *p, vreal *Ap, const vreal *r0star,
                       vreal *ApAp, vreal *pAp, vreal *Apr0)
{
  vreal s_ApAp = 0, s_pAp = 0, s_Apr0 = 0;
#ifdef _OPENMP
  #pragma omp parallel for reduction(+: s_ApAp, s_pAp, s_Apr0)
#endif
  for (long row = 0; row < (long)rows; ++row)
  {
    vreal dot = 0;
    vreal pd = p[row];
    size_t row_end = rp[row + 1];
    for (size_t i = rp[row]; i < row_end; ++i)
      dot += v[i] * p[ci[i]];
    Ap[row] = dot;
    s_ApAp += dot * dot;
    s_pAp  += pd * dot;
    s_Apr0 += r0star ? dot * r0star[row] : 0.0;
  }
  *ApAp = s_ApAp; *pAp = s_pAp;
  if (r0star && Apr0) *Apr0 = s_Apr0;
}

#define HIST_PUSH(val) do { if (hist_len) { if (hist && *hist_len < hist_cap) hist[*hist_len] = (val); (*hist_len)++; } } while (0)

/* ------------------------------------------------------------------------------------------------
 * Pipelined CG (Chronopoulos/Gear): viennacl/linalg/cg.hpp:128-187 with
 * host_based/iterative_operations.hpp:378-418 (vector update) and :58-103 (SpMV + dots).
 * ---------------------------------------------------------------------------------------------- */
int vclo_cg(int rows, const u32 *rp, const u32 *ci, const vreal *v,
            const vreal *b, vreal *x, double tol, double abs_tol, int maxit,
            int *iters, double *err, double *hist, int hist_cap, int *hist_len)
{
  size_t n = (size_t)rows;
  vreal *r = (vreal*)malloc(sizeof(vreal) * n), *p = (vreal*)malloc(sizeof(vreal) * n), *Ap = (vreal*)malloc(sizeof(vreal) * n);
  if (hist_len) *hist_len = 0;
  memset(x, 0, sizeof(vreal) * n);
  memcpy(r, b, sizeof(vreal) * n);
  memcpy(p, b, sizeof(vreal) * n);
  vclo_csr_spmv(rows, rp, ci, v, p, 0, 1, 1.0, Ap, 0, 1, 0.0);

  vreal norm_rhs_squared = vclo_norm2(r, rows); norm_rhs_squared *= norm_rhs_squared;
  *iters = 0; *err = 0;
  if (norm_rhs_squared <= abs_tol * abs_tol) { free(r); free(p); free(Ap); return 0; }

  vreal rr = norm_rhs_squared;
  vreal alpha = rr / vclo_inner_prod(p, Ap, rows);
  vreal beta = vclo_norm2(Ap, rows); beta = (alpha * alpha * beta * beta - rr) / rr;
  vreal ApAp = 0, pAp = 0;

  for (int i = 0; i < maxit; ++i)
  {
    *iters = i + 1;
    vreal s_rr = 0;
#ifdef _OPENMP
    #pragma omp parallel for reduction(+: s_rr)
#endif
    for (long k = 0; k < (long)rows; ++k)
    {
      vreal vp = p[k], vr = r[k];
      x[k] += alpha * vp;
      vr -= alpha * Ap[k];
      vp = vr + beta * vp;
      s_rr += vr * vr;
      p[k] = vp; r[k] = vr;
    }
    rr = s_rr;
    fused_prod(rows, rp, ci, v, p, Ap, NULL, &ApAp, &pAp, NULL);

    HIST_PUSH(sqrt(fabs(rr / norm_rhs_squared)));
    if (fabs(rr / norm_rhs_squared) < tol * tol || fabs(rr) < abs_tol * abs_tol) break;

    alpha = rr / pAp;
    beta = (alpha * alpha * ApAp - rr) / rr;
  }
  *err = sqrt(fabs(rr) / norm_rhs_squared);
  free(r); free(p); free(Ap);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Pipelined BiCGStab: viennacl/linalg/bicgstab.hpp:97-215 with host_based/iterative_operations.hpp:518-621.
 * NB (reference behaviour, kept): on convergence the loop breaks BEFORE the final vector update, so the returned
 * x is the iterate of the previous step while tag.error() describes the new residual.
 * ---------------------------------------------------------------------------------------------- */
int vclo_bicgstab(int rows, const u32 *rp, const u32 *ci, const vreal *v,
                  const vreal *b, vreal *x, double tol, double abs_tol, int maxit,
                  int *iters, double *err, double *hist, int hist_cap, int *hist_len)
{
  size_t n = (size_t)rows;
  vreal *r = (vreal*)malloc(sizeof(vreal) * n), *p = (vreal*)malloc(sizeof(vreal) * n), *r0 = (vreal*)malloc(sizeof(vreal) * n);
  vreal *Ap = (vreal*)malloc(sizeof(vreal) * n), *s = (vreal*)malloc(sizeof(vreal) * n), *As = (vreal*)malloc(sizeof(vreal) * n);
  if (hist_len) *hist_len = 0;
  memset(x, 0, sizeof(vreal) * n);
  memcpy(r, b, sizeof(vreal) * n); memcpy(p, b, sizeof(vreal) * n); memcpy(r0, b, sizeof(vreal) * n);
  memcpy(Ap, b, sizeof(vreal) * n); memcpy(s, b, sizeof(vreal) * n); memcpy(As, b, sizeof(vreal) * n);

  vreal norm_rhs = vclo_norm2(r, rows);
  vreal residual_norm = norm_rhs;
  vreal r_dot_r0 = norm_rhs * norm_rhs;            /* inner_prod_buffer[0] = ||b||^2 (bicgstab.hpp:131) */
  vreal As_As = 0, As_s = 0, Ap_r0 = 0, As_r0 = 0, s_s = 0, dummy1, dummy2;
  *iters = 0; *err = 0;
  if (norm_rhs <= abs_tol) goto done;

  for (int i = 0; i < maxit; ++i)
  {
    *iters = i + 1;
    fused_prod(rows, rp, ci, v, p, Ap, r0, &dummy1, &dummy2, &Ap_r0);

    /* update_s: alpha on "device" from chunks 0 and 3 */
    vreal alpha = r_dot_r0 / Ap_r0;
    vreal t_ss = 0;
#ifdef _OPENMP
    #pragma omp parallel for reduction(+: t_ss)
#endif
    for (long k = 0; k < (long)rows; ++k)
    {
      vreal vs = r[k] - alpha * Ap[k];
      t_ss += vs * vs;
      s[k] = vs;
    }
    s_s = t_ss;

    fused_prod(rows, rp, ci, v, s, As, r0, &As_As, &As_s, &As_r0);

    alpha = r_dot_r0 / Ap_r0;
    vreal beta = -As_r0 / Ap_r0;
    vreal omega = As_s / As_As;

    residual_norm = sqrt(s_s - 2.0 * omega * As_s + omega * omega * As_As);
    HIST_PUSH(fabs(residual_norm / norm_rhs));
    if (fabs(residual_norm / norm_rhs) < tol || residual_norm < abs_tol) break;

    vreal t_rr0 = 0;
#ifdef _OPENMP
    #pragma omp parallel for reduction(+: t_rr0)
#endif
    for (long k = 0; k < (long)rows; ++k)
    {
      vreal vx = x[k], vp = p[k], vs = s[k], vr, vAs = As[k], vAp = Ap[k];
      vx += alpha * vp + omega * vs;
      vr  = vs - omega * vAs;
      vp  = vr + beta * (vp - omega * vAp);
      t_rr0 += vr * r0[k];
      x[k] = vx; r[k] = vr; p[k] = vp;
    }
    r_dot_r0 = t_rr0;
  }
  *err = residual_norm / norm_rhs;
done:
  free(r); free(p); free(r0); free(Ap); free(s); free(As);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Preconditioned (left) BiCGStab, generic path: viennacl/linalg/bicgstab.hpp:398-489;
 * Jacobi: jacobi_precond.hpp:103-130 (vec = element_div(vec, diag)).
 * ---------------------------------------------------------------------------------------------- */
static void apply_precond(int precond, const vreal *diag, vreal *vec, int rows)
{
  if (precond != 1) return;
#ifdef _OPENMP
  #pragma omp parallel for
#endif
  for (long k = 0; k < (long)rows; ++k) vec[k] = vec[k] / diag[k];
}

int vclo_bicgstab_precond(int rows, const u32 *rp, const u32 *ci, const vreal *v, int precond,
                          const vreal *b, vreal *x, double tol, double abs_tol, int maxit, int restart_every,
                          int *iters, double *err, double *hist, int hist_cap, int *hist_len)
{
  size_t n = (size_t)rows;
  vreal *r = (vreal*)malloc(sizeof(vreal) * n), *p = (vreal*)malloc(sizeof(vreal) * n), *r0 = (vreal*)malloc(sizeof(vreal) * n);
  vreal *t0 = (vreal*)malloc(sizeof(vreal) * n), *t1 = (vreal*)malloc(sizeof(vreal) * n), *s = (vreal*)malloc(sizeof(vreal) * n);
  vreal *diag = (vreal*)malloc(sizeof(vreal) * n);
  if (hist_len) *hist_len = 0;
  vclo_csr_diag(rows, rp, ci, v, diag);
  memset(x, 0, sizeof(vreal) * n);
  memcpy(r, b, sizeof(vreal) * n); memcpy(p, b, sizeof(vreal) * n); memcpy(r0, b, sizeof(vreal) * n);

  vreal ip_rr0 = vclo_norm2(r, rows);
  vreal norm_rhs = vclo_norm2(r, rows);
  vreal residual_norm = norm_rhs, new_ip_rr0 = 0;
  *iters = 0; *err = 0;
  if (norm_rhs <= abs_tol) goto done;
  {
    int restart_flag = 1;
    long last_restart = 0;
    for (long i = 0; i < maxit; ++i)
    {
      if (restart_flag)
      {
        vclo_csr_spmv(rows, rp, ci, v, x, 0, 1, 1.0, r, 0, 1, 0.0);
        for (size_t k = 0; k < n; ++k) r[k] = b[k] - r[k];
        apply_precond(precond, diag, r, rows);
        memcpy(p, r, sizeof(vreal) * n); memcpy(r0, r, sizeof(vreal) * n);
        ip_rr0 = vclo_norm2(r, rows); ip_rr0 *= ip_rr0;
        restart_flag = 0; last_restart = i;
      }
      *iters = (int)(i + 1);
      vclo_csr_spmv(rows, rp, ci, v, p, 0, 1, 1.0, t0, 0, 1, 0.0);
      apply_precond(precond, diag, t0, rows);
      vreal alpha = ip_rr0 / vclo_inner_prod(t0, r0, rows);
      for (size_t k = 0; k < n; ++k) s[k] = r[k] - alpha * t0[k];

      vclo_csr_spmv(rows, rp, ci, v, s, 0, 1, 1.0, t1, 0, 1, 0.0);
      apply_precond(precond, diag, t1, rows);
      vreal norm_t1 = vclo_norm2(t1, rows);
      vreal omega = vclo_inner_prod(t1, s, rows) / (norm_t1 * norm_t1);

      for (size_t k = 0; k < n; ++k) x[k] += alpha * p[k] + omega * s[k];
      for (size_t k = 0; k < n; ++k) r[k] = s[k] - omega * t1[k];

      residual_norm = vclo_norm2(r, rows);
      HIST_PUSH(fabs(residual_norm / norm_rhs));
      if (residual_norm / norm_rhs < tol || residual_norm < abs_tol) break;

      new_ip_rr0 = vclo_inner_prod(r, r0, rows);
      vreal beta = new_ip_rr0 / ip_rr0 * alpha / omega;
      ip_rr0 = new_ip_rr0;

      if ((ip_rr0 >= 0 && ip_rr0 <= 0) || (omega >= 0 && omega <= 0) || i - last_restart > restart_every)
        restart_flag = 1;

      for (size_t k = 0; k < n; ++k) p[k] -= omega * t0[k];
      for (size_t k = 0; k < n; ++k) p[k] = r[k] + beta * p[k];
    }
    *err = residual_norm / norm_rhs;
  }
done:
  free(r); free(p); free(r0); free(t0); free(t1); free(s); free(diag);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Pipelined "simpler GMRES" with classical Gram-Schmidt: viennacl/linalg/gmres.hpp:181-367, fused steps with the
 * semantics of cuda/iterative_operations.hpp:1597-1894 (== host_based/iterative_operations.hpp:733-937 minus the
 * dropped-tail defect at :820-822).  Basis vector k lives at k*internal_size (size padded to 128, gmres.hpp:192).
 * ---------------------------------------------------------------------------------------------- */
int vclo_gmres(int rows, const u32 *rp, const u32 *ci, const vreal *v,
               const vreal *b, vreal *x, double tol, double abs_tol, int maxit, int krylov,
               int *iters, double *err, double *hist, int hist_cap, int *hist_len)
{
  size_t n = (size_t)rows;
  size_t isz = (n + 127) / 128 * 128;
  size_t m = (size_t)krylov;
  vreal *res = (vreal*)malloc(sizeof(vreal) * n);
  vreal *V = (vreal*)calloc(isz * m, sizeof(vreal));
  vreal *R = (vreal*)calloc(m * m, sizeof(vreal));
  vreal *xi = (vreal*)calloc(m, sizeof(vreal)), *eta = (vreal*)calloc(m, sizeof(vreal)), *coef = (vreal*)calloc(m, sizeof(vreal));
  vreal *h = (vreal*)calloc(m, sizeof(vreal));
  if (hist_len) *hist_len = 0;
  memset(x, 0, sizeof(vreal) * n);
  memcpy(res, b, sizeof(vreal) * n);

  vreal norm_rhs = vclo_norm2(res, rows);
  vreal rho_0 = norm_rhs, rho = 1.0;
  *iters = 0; *err = 0;

  unsigned max_restarts = (unsigned)maxit / (unsigned)krylov;                  /* gmres.hpp:74-80 */
  if (max_restarts > 0 && max_restarts * (unsigned)krylov == (unsigned)maxit) max_restarts -= 1;

  for (unsigned restart = 0; restart <= max_restarts; ++restart)
  {
    if (restart > 0)
    {
      vclo_csr_spmv(rows, rp, ci, v, x, 0, 1, 1.0, res, 0, 1, 0.0);
      for (size_t k = 0; k < n; ++k) res[k] = b[k] - res[k];
      rho_0 = vclo_norm2(res, rows);
    }
    if (rho_0 <= abs_tol) break;
    for (size_t k = 0; k < n; ++k) res[k] /= rho_0;
    rho = 1.0;
    if (rho_0 / norm_rhs < tol || rho_0 < abs_tol) break;

    size_t k;
    for (k = 0; k < m; ++k)
    {
      vreal *vk = V + k * isz;
      const vreal *src = (k == 0) ? res : V + (k - 1) * isz;
      vreal ApAp, pAp;
      fused_prod(rows, rp, ci, v, src, vk, NULL, &ApAp, &pAp, NULL);   /* chunk 1 <- <v_k,v_k> */
      vreal norm_sq = ApAp;
      if (k > 0)
      {
        /* stage 1: h_j = <v_j, v_k>, j < k */
        for (size_t j = 0; j < k; ++j) h[j] = vclo_inner_prod(V + j * isz, vk, rows);
        /* stage 2: v_k -= sum h_j v_j ; R[j + k*m] = h_j ; ||v_k||^2 */
        vreal nsq = 0;
#ifdef _OPENMP
        #pragma omp parallel for reduction(+: nsq)
#endif
        for (long i = 0; i < (long)rows; ++i)
        {
          vreal val = vk[i];
          for (size_t j = 0; j < k; ++j) val -= h[j] * V[(size_t)i + j * isz];
          nsq += val * val;
          vk[i] = val;
        }
        for (size_t j = 0; j < k; ++j) R[j + k * m] = h[j];
        norm_sq = nsq;
      }
      /* normalize: R[k + k*m] = ||v_k||, v_k /= ||v_k||, xi_k = <r, v_k> */
      vreal nrm = sqrt(norm_sq);
      R[k + k * m] = nrm;
      vreal rv = 0;
#ifdef _OPENMP
      #pragma omp parallel for reduction(+: rv)
#endif
      for (long i = 0; i < (long)rows; ++i)
      {
        vreal val = vk[i] / nrm;
        rv += res[i] * val;
        vk[i] = val;
      }
      xi[k] = rv;
    }

    /* premature convergence check (gmres.hpp:306-314): note the reference indexes R[i + i*k] with the current k */
    size_t full = k;
    for (size_t i = 0; i < k; ++i)
      if (fabs(R[i + i * k]) < tol * R[0]) { k = i; break; }

    for (size_t i = 0; i < k; ++i)
    {
      *iters += 1;
      if (xi[i] >= rho || xi[i] <= -rho) { k = i; break; }
      rho *= sin(acos(xi[i] / rho));
    }

    memcpy(eta, xi, sizeof(vreal) * m);
    for (long i2 = (long)k - 1; i2 > -1; --i2)
    {
      size_t i = (size_t)i2;
      for (size_t j = i + 1; j < k; ++j) eta[i] -= R[i + j * full] * eta[j];
      eta[i] /= R[i + i * full];
    }
    for (size_t i = 0; i < k; ++i) coef[i] = rho_0 * eta[i];

    /* x += c_0 r + sum_{j>=1} c_j v_{j-1}  (host_based/iterative_operations.hpp:895-922) */
    /* (called even for k == 0, with whatever coef[0] holds from the previous cycle -- reference behaviour) */
#ifdef _OPENMP
    #pragma omp parallel for
#endif
    for (long i = 0; i < (long)rows; ++i)
    {
      vreal val = x[i];
      val += coef[0] * res[i];
      for (size_t j = 1; j < k; ++j) val += coef[j] * V[(size_t)i + (j - 1) * isz];
      x[i] = val;
    }
    *err = fabs(rho * rho_0 / norm_rhs);
    HIST_PUSH(fabs(rho * rho_0 / norm_rhs));
  }
  free(res); free(V); free(R); free(xi); free(eta); free(coef); free(h);
  return 0;
}
