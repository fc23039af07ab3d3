/* vcl_b200_float.h -- single-precision half of the C-ABI of libvcl_b200.so.  GENERATED from vcl_b200.h by
 * tools/gen_float_header.py (do not edit): the same entry points with the precision letter S, float vectors,
 * matrices and scalars (reference: every type on the path is a template over NumericT = float | double).
 * The row-partitioned (multi-GPU) path exists in double precision only. */
#ifndef VCL_B200_FLOAT_H
#define VCL_B200_FLOAT_H

#include "vcl_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- BLAS-1 subset ----------------------------------------------------------- */
/* linalg/cuda/vector_operations.hpp:77 (av), :179 (avbv), :483 (avbv_v), :782 (vector_assign), :870 (element_op /),
 * :1273-1579 (inner_prod + inner_prod_cpu), :2018-2448 (norm_2 + norm_2_cpu). */
ViennaCLStatus ViennaCLCUDASav(ViennaCLBackend backend, ViennaCLInt n, float *x, ViennaCLInt offx, ViennaCLInt incx,
                               const float *y, ViennaCLInt offy, ViennaCLInt incy, float alpha);                     /* x = alpha*y */
ViennaCLStatus ViennaCLCUDASavbv(ViennaCLBackend backend, ViennaCLInt n, float *x, ViennaCLInt offx, ViennaCLInt incx,
                                 const float *y, ViennaCLInt offy, ViennaCLInt incy, float alpha,
                                 const float *z, ViennaCLInt offz, ViennaCLInt incz, float beta);                    /* x = alpha*y + beta*z */
ViennaCLStatus ViennaCLCUDASavbv_v(ViennaCLBackend backend, ViennaCLInt n, float *x, ViennaCLInt offx, ViennaCLInt incx,
                                   const float *y, ViennaCLInt offy, ViennaCLInt incy, float alpha,
                                   const float *z, ViennaCLInt offz, ViennaCLInt incz, float beta);                  /* x += alpha*y + beta*z */
ViennaCLStatus ViennaCLCUDASassign(ViennaCLBackend backend, ViennaCLInt n, float *x, ViennaCLInt offx, ViennaCLInt incx, float value);
ViennaCLStatus ViennaCLCUDASelement_div(ViennaCLBackend backend, ViennaCLInt n, float *x, ViennaCLInt offx, ViennaCLInt incx,
                                        const float *y, ViennaCLInt offy, ViennaCLInt incy,
                                        const float *z, ViennaCLInt offz, ViennaCLInt incz);                          /* x = y ./ z */
ViennaCLStatus ViennaCLCUDASdot(ViennaCLBackend backend, ViennaCLInt n, float *result_host,
                                const float *x, ViennaCLInt offx, ViennaCLInt incx,
                                const float *y, ViennaCLInt offy, ViennaCLInt incy);                                  /* synchronous */
ViennaCLStatus ViennaCLCUDASnrm2(ViennaCLBackend backend, ViennaCLInt n, float *result_host,
                                 const float *x, ViennaCLInt offx, ViennaCLInt incx);                                 /* synchronous */

/* ---------------------------------------------------------------- SpMV -------------------------------------------------------------------- */
/* CSR row blocks: compressed_matrix.hpp:1152-1188 (generate_row_block_information) -> handle3()/blocks1().
 * Each block holds whole rows: at most VCL_B200_CSR_BLOCK_ROWS rows and VCL_B200_CSR_BLOCK_NNZ non-zeros, or one longer row.
 * Two-call protocol: row_blocks == NULL returns the count in *num_blocks; then pass a device buffer of (*num_blocks + 1) u32. */
#ifndef VCL_B200_CSR_BLOCK_ROWS
#define VCL_B200_CSR_BLOCK_ROWS 256
#endif
#ifndef VCL_B200_CSR_BLOCK_NNZ
#define VCL_B200_CSR_BLOCK_NNZ  2048
#endif

/* y[offy + i*incy] = alpha * (A x)_i + (beta != 0 ? beta * y[...] : 0);  x read at offx + col*incx.
 * linalg/sparse_matrix_operations.hpp:90-121 -> cuda/sparse_matrix_operations.hpp:262-396 (kernels :137-249).
 * row_blocks: a plan from ViennaCLCUDAcsr_row_blocks, or ANY partition of the rows into consecutive blocks (e.g. the reference's own
 * handle3() plan) -- a plan this library did not make is checked once per plan address (cached; rewriting or freeing the plan's
 * memory through this API drops the verdict) and, when a block breaks the limits above, the product runs through the plan-free
 * kernel (one thread per row, no staging) -- as it does for row_blocks == NULL or arrays that are not 16-byte aligned. */
ViennaCLStatus ViennaCLCUDAScsrmv(ViennaCLBackend backend, ViennaCLInt rows, ViennaCLInt cols, ViennaCLInt nnz,
                                  const unsigned int *row_ptr, const unsigned int *col_idx, const float *values,
                                  const unsigned int *row_blocks, ViennaCLInt num_blocks,
                                  const float *x, ViennaCLInt offx, ViennaCLInt incx, float alpha,
                                  float *y, ViennaCLInt offy, ViennaCLInt incy, float beta);

/* SELL-C-sigma (sigma = 1): cuda/sparse_matrix_operations.hpp:2196-2289; layout sliced_ell_matrix.hpp:134-214. */
ViennaCLStatus ViennaCLCUDASsellmv(ViennaCLBackend backend, ViennaCLInt rows, ViennaCLInt cols, ViennaCLInt rows_per_block,
                                   const unsigned int *columns_per_block, const unsigned int *col_idx,
                                   const unsigned int *block_start, const float *values,
                                   const float *x, ViennaCLInt offx, ViennaCLInt incx, float alpha,
                                   float *y, ViennaCLInt offy, ViennaCLInt incy, float beta);

/* Device-side CSR -> SELL-C conversion with the exact array layout of sliced_ell_matrix.hpp:140-214.
 * Call 1: columns_per_block/block_start sized ceil(rows/C) are filled and *padded_nnz returned (col_idx/values NULL).
 * Call 2: col_idx/values sized *padded_nnz are filled. */
ViennaCLStatus ViennaCLCUDAScsr2sell(ViennaCLBackend backend, ViennaCLInt rows, ViennaCLInt rows_per_block,
                                     const unsigned int *row_ptr, const unsigned int *csr_col, const float *csr_val,
                                     unsigned int *columns_per_block, unsigned int *block_start, long long *padded_nnz,
                                     unsigned int *col_idx, float *values);
/* SELL-C-sigma (not in the reference, whose sigma is fixed at 1): inside every window of `sigma` consecutive rows
 * (sigma a multiple of rows_per_block, <= 4096) the rows are ordered by decreasing length -- stable, ties keep their order --
 * before they are cut into slices, which removes most of the padding of irregular matrices.  row_perm (out, ceil(rows/C)*C
 * entries) maps storage rows to matrix rows.  Same two-call protocol as csr2sell: the first call (col_idx == NULL) fills
 * row_perm, columns_per_block, block_start and *padded_nnz, the second one scatters the entries.  Products and solvers take
 * the matrix through the ViennaCLCUDASsell struct (row_perm set); per-row arithmetic and results equal those of sigma = 1. */
ViennaCLStatus ViennaCLCUDAScsr2sell_sigma(ViennaCLBackend backend, ViennaCLInt rows, ViennaCLInt rows_per_block, ViennaCLInt sigma,
                                           const unsigned int *row_ptr, const unsigned int *csr_col, const float *csr_val,
                                           unsigned int *row_perm, unsigned int *columns_per_block, unsigned int *block_start,
                                           long long *padded_nnz, unsigned int *col_idx, float *values);

/* ELL (ell_matrix.hpp:36-119) and HYB (hyb_matrix.hpp:36-126), AlignmentV = 1 layouts:
 * ELL entry j of row r at j*internal_rows + r (coords / elements hold internal_rows*maxnnz entries, padding value 0, column 0);
 * HYB = ELL part of width ell.maxnnz + CSR tail (csr_rows[rows+1], csr_cols, csr_elements).
 * Products: cuda/sparse_matrix_operations.hpp:1747-1838 (ELL), :2298-2400 (HYB); zero-valued ELL slots never touch x. */
typedef struct
{
  ViennaCLInt rows, cols, internal_rows, maxnnz;
  const unsigned int *coords;
  const float *elements;
} ViennaCLCUDASell;

typedef struct
{
  ViennaCLCUDASell ell;
  const unsigned int *csr_rows, *csr_cols;
  const float *csr_elements;
  ViennaCLInt csr_nnz;
} ViennaCLCUDAShyb;

ViennaCLStatus ViennaCLCUDASellmv(ViennaCLBackend backend, const ViennaCLCUDASell *A,
                                  const float *x, ViennaCLInt offx, ViennaCLInt incx, float alpha,
                                  float *y, ViennaCLInt offy, ViennaCLInt incy, float beta);
ViennaCLStatus ViennaCLCUDAShybmv(ViennaCLBackend backend, const ViennaCLCUDAShyb *A,
                                  const float *x, ViennaCLInt offx, ViennaCLInt incx, float alpha,
                                  float *y, ViennaCLInt offy, ViennaCLInt incy, float beta);
/* Device-side CSR -> ELL (layout of ell_matrix.hpp:122-166).  Call 1 (coords == NULL): *maxnnz = longest row.
 * Call 2: coords / elements sized rows * (*maxnnz) are filled (internal_rows = rows). */
ViennaCLStatus ViennaCLCUDAScsr2ell(ViennaCLBackend backend, ViennaCLInt rows, const unsigned int *row_ptr,
                                    const unsigned int *csr_col, const float *csr_val, ViennaCLInt *maxnnz,
                                    unsigned int *coords, float *elements);
/* Device-side CSR -> HYB (layout and width rule of hyb_matrix.hpp:127-214: the smallest width that covers at least
 * `csr_threshold` (reference default 0.8) of the rows).  Call 1 (ell_coords == NULL): *ell_width and *csr_nnz (>= 1: the
 * reference stores one dummy entry when the tail is empty).  Call 2 fills all five arrays. */
ViennaCLStatus ViennaCLCUDAScsr2hyb(ViennaCLBackend backend, ViennaCLInt rows, ViennaCLInt cols, const unsigned int *row_ptr,
                                    const unsigned int *csr_col, const float *csr_val, float csr_threshold,
                                    ViennaCLInt *ell_width, ViennaCLInt *csr_nnz,
                                    unsigned int *ell_coords, float *ell_elements,
                                    unsigned int *csr_rows, unsigned int *csr_cols, float *csr_elements);

/* COO (coordinate_matrix.hpp:47-102): coords = (row, col) pairs, entries sorted by row as the reference's copy() produces
 * them.  SpMV and the solvers run on a CSR index of the same entries, built once on the device:
 * row_ptr[rows+1] and col_idx[nnz] are filled, the value array is shared with the COO matrix (no copy).
 * Fails with ViennaCLB200InvalidArgument when the entries are not sorted by row. */
/* coordinate_matrix product with the reference's arithmetic (host_based/sparse_matrix_operations.hpp:1222-1247):
 * y <- beta*y (or 0), then y[row] += (alpha*a) * x[col] entry by entry; (row_ptr, col_idx) from ViennaCLCUDAcoo2csr. */
ViennaCLStatus ViennaCLCUDAScoomv(ViennaCLBackend backend, ViennaCLInt rows, ViennaCLInt cols, ViennaCLInt nnz,
                                  const unsigned int *row_ptr, const unsigned int *col_idx, const float *elements,
                                  const unsigned int *row_blocks, ViennaCLInt num_blocks,
                                  const float *x, ViennaCLInt offx, ViennaCLInt incx, float alpha,
                                  float *y, ViennaCLInt offy, ViennaCLInt incy, float beta);

/* detail::row_info: linalg/sparse_matrix_operations.hpp:48-74 -> cuda/sparse_matrix_operations.hpp:53-119.
 * option: 0 inf-norm, 1 1-norm, 2 2-norm, 3 diagonal (forwards.h row_info_types order). */
ViennaCLStatus ViennaCLCUDAScsr_row_info(ViennaCLBackend backend, ViennaCLInt rows,
                                         const unsigned int *row_ptr, const unsigned int *col_idx, const float *values,
                                         float *result, ViennaCLInt option);

/* ---------------------------------------------------------------- matrix / vector generators ---------------------------------------------- */
/* tools/matrix_generation.hpp:47-88 generalised (DESIGN.md "synthetic inputs"): 5-/7-point stencil with first-order upwind
 * convection c; nz == 1 selects the 2-D 5-point stencil.  row_ptr[rows+1], col_idx[nnz], values[nnz] are device buffers;
 * pass NULLs to get the counts only. */
ViennaCLStatus ViennaCLCUDASgenerate_stencil(ViennaCLBackend backend, ViennaCLInt nx, ViennaCLInt ny, ViennaCLInt nz,
                                             float cx, float cy, float cz,
                                             unsigned int *row_ptr, unsigned int *col_idx, float *values,
                                             long long *rows, long long *nnz);
/* Same matrix restricted to rows [row_begin, row_end) (global column indices, local row_ptr starting at 0). */
ViennaCLStatus ViennaCLCUDASgenerate_stencil_rows(ViennaCLBackend backend, ViennaCLInt nx, ViennaCLInt ny, ViennaCLInt nz,
                                                  float cx, float cy, float cz, long long row_begin, long long row_end,
                                                  unsigned int *row_ptr, unsigned int *col_idx, float *values, long long *nnz);
ViennaCLStatus ViennaCLCUDASfill_uniform(ViennaCLBackend backend, long long n, float *x, unsigned long long seed,
                                         long long index_offset, float lo, float hi);

/* ---------------------------------------------------------------- fused solver steps ------------------------------------------------------ */
/* One-to-one with linalg/iterative_operations.hpp (argument lists :59-65, :97-100, :134-139, :171-176, :208-214, :248-255,
 * :286-291, :321-329, :356-362, :393-396).  `buf` is the reference's inner_prod_buffer: chunks of `chunk` entries; each
 * routine writes the FULLY REDUCED value into element 0 of its chunk(s) and leaves the rest untouched (the reference's
 * host backend does the same, host_based/iterative_operations.hpp:100-102; the drivers sum whole chunks). */
typedef struct
{
  ViennaCLInt rows, cols, nnz;
  const unsigned int *row_ptr, *col_idx;
  const float *values;
  const unsigned int *row_blocks;   /* may be NULL */
  ViennaCLInt num_blocks;
} ViennaCLCUDAScsr;

typedef struct
{
  ViennaCLInt rows, cols, rows_per_block;
  const unsigned int *columns_per_block, *col_idx, *block_start;
  const float *values;
  const unsigned int *row_perm;   /* NULL: sigma = 1, the reference's layout (sliced_ell_matrix.hpp:43).  Otherwise SELL-C-sigma:
                                     storage row i holds matrix row row_perm[i] (0xFFFFFFFF: padding row), see csr2sell_sigma */
} ViennaCLCUDASsell;
/* y = alpha*A*x + beta*y for a SELL matrix passed as the struct (needed for SELL-C-sigma: row_perm); ViennaCLCUDASsellmv is the
 * flat-argument form of the reference's kernel signature. */
ViennaCLStatus ViennaCLCUDASsellmv_struct(ViennaCLBackend backend, const ViennaCLCUDASsell *A,
                                          const float *x, ViennaCLInt offx, ViennaCLInt incx, float alpha,
                                          float *y, ViennaCLInt offy, ViennaCLInt incy, float beta);

ViennaCLStatus ViennaCLCUDASpipelined_cg_vector_update(ViennaCLBackend backend, ViennaCLInt n, float *result, float alpha,
                                                       float *p, float *r, const float *Ap, float beta,
                                                       float *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDASpipelined_cg_prod_csr(ViennaCLBackend backend, const ViennaCLCUDAScsr *A, const float *p, float *Ap,
                                                  float *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDASpipelined_cg_prod_sell(ViennaCLBackend backend, const ViennaCLCUDASsell *A, const float *p, float *Ap,
                                                   float *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDASpipelined_bicgstab_update_s(ViennaCLBackend backend, ViennaCLInt n, float *s, const float *r, const float *Ap,
                                                        float *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset);
ViennaCLStatus ViennaCLCUDASpipelined_bicgstab_vector_update(ViennaCLBackend backend, ViennaCLInt n, float *result, float alpha, float *p,
                                                             float omega, const float *s, float *residual, const float *As,
                                                             float beta, const float *Ap, const float *r0star,
                                                             float *buf, ViennaCLInt chunk);
ViennaCLStatus ViennaCLCUDASpipelined_bicgstab_prod_csr(ViennaCLBackend backend, const ViennaCLCUDAScsr *A, const float *p, float *Ap,
                                                        const float *r0star, float *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset);
ViennaCLStatus ViennaCLCUDASpipelined_bicgstab_prod_sell(ViennaCLBackend backend, const ViennaCLCUDASsell *A, const float *p, float *Ap,
                                                         const float *r0star, float *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset);
ViennaCLStatus ViennaCLCUDASpipelined_gmres_normalize_vk(ViennaCLBackend backend, ViennaCLInt n, float *v_k, const float *residual,
                                                         float *R, ViennaCLInt offset_in_R, const float *buf,
                                                         float *r_dot_vk, ViennaCLInt chunk, ViennaCLInt chunk_offset);
ViennaCLStatus ViennaCLCUDASpipelined_gmres_gram_schmidt_stage1(ViennaCLBackend backend, const float *basis, ViennaCLInt n,
                                                                ViennaCLInt internal_n, ViennaCLInt k, float *vi_in_vk, ViennaCLInt chunk);
ViennaCLStatus ViennaCLCUDASpipelined_gmres_gram_schmidt_stage2(ViennaCLBackend backend, float *basis, ViennaCLInt n,
                                                                ViennaCLInt internal_n, ViennaCLInt k, const float *vi_in_vk,
                                                                float *R, ViennaCLInt krylov_dim, float *buf, ViennaCLInt chunk);
ViennaCLStatus ViennaCLCUDASpipelined_gmres_update_result(ViennaCLBackend backend, ViennaCLInt n, float *result, const float *residual,
                                                          const float *basis, ViennaCLInt internal_n, const float *coefficients, ViennaCLInt k);
ViennaCLStatus ViennaCLCUDASpipelined_gmres_prod_csr(ViennaCLBackend backend, const ViennaCLCUDAScsr *A, const float *p, float *Ap,
                                                     float *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDASpipelined_gmres_prod_sell(ViennaCLBackend backend, const ViennaCLCUDASsell *A, const float *p, float *Ap,
                                                      float *buf, ViennaCLInt buf_size);
/* the same fused products for ell_matrix / hyb_matrix (cuda/iterative_operations.hpp:330-727, :1138-1593) */
ViennaCLStatus ViennaCLCUDASpipelined_cg_prod_ell(ViennaCLBackend backend, const ViennaCLCUDASell *A, const float *p, float *Ap,
                                                  float *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDASpipelined_cg_prod_hyb(ViennaCLBackend backend, const ViennaCLCUDAShyb *A, const float *p, float *Ap,
                                                  float *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDASpipelined_bicgstab_prod_ell(ViennaCLBackend backend, const ViennaCLCUDASell *A, const float *p, float *Ap,
                                                        const float *r0star, float *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset);
ViennaCLStatus ViennaCLCUDASpipelined_bicgstab_prod_hyb(ViennaCLBackend backend, const ViennaCLCUDAShyb *A, const float *p, float *Ap,
                                                        const float *r0star, float *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset);
ViennaCLStatus ViennaCLCUDASpipelined_gmres_prod_ell(ViennaCLBackend backend, const ViennaCLCUDASell *A, const float *p, float *Ap,
                                                     float *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDASpipelined_gmres_prod_hyb(ViennaCLBackend backend, const ViennaCLCUDAShyb *A, const float *p, float *Ap,
                                                     float *buf, ViennaCLInt buf_size);

/* ---------------------------------------------------------------- whole solves ------------------------------------------------------------ */
/* The loop lives next to the kernels (device-resident scalars, no per-iteration host round trip); the C++ `solve()` keeps its
 * signature and calls these.  Semantics (tolerances, iteration counting, error estimate, quirks) are those of
 * linalg/cg.hpp:128-187, bicgstab.hpp:97-215 / :398-489 and gmres.hpp:181-367.
 * monitor (optional) has the reference's contract: called with the device pointer of the current iterate and the relative
 * residual estimate, once per iteration (GMRES: once per restart); returning non-zero stops the solver. */
typedef ViennaCLInt (*ViennaCLMonitorS)(const float *x_dev, float rel_residual_estimate, void *user);

/* Diagonal preconditioners that the drivers fold into their kernels (CSR matrices): Jacobi (jacobi_precond.hpp:103-130,
 * divide by diag(A)) and row scaling (row_scaling.hpp:150-190, divide by the inf-/1-/2-norm of the row). */

typedef struct
{
  double tolerance;                 /* relative */
  double abs_tolerance;
  ViennaCLInt max_iterations;
  ViennaCLInt krylov_dim;           /* GMRES only */
  ViennaCLInt max_iterations_before_restart; /* BiCGStab (preconditioned path) only */
  ViennaCLB200Precond precond;
  ViennaCLMonitorS monitor;
  void *monitor_user;
  /* results */
  ViennaCLInt iters;
  double error;
} ViennaCLB200SolverTagS;

ViennaCLStatus ViennaCLCUDAScsr_cg(ViennaCLBackend backend, const ViennaCLCUDAScsr *A, const float *b, float *x, ViennaCLB200SolverTagS *tag);
ViennaCLStatus ViennaCLCUDASsell_cg(ViennaCLBackend backend, const ViennaCLCUDASsell *A, const float *b, float *x, ViennaCLB200SolverTagS *tag);
ViennaCLStatus ViennaCLCUDAScsr_bicgstab(ViennaCLBackend backend, const ViennaCLCUDAScsr *A, const float *b, float *x, ViennaCLB200SolverTagS *tag);
ViennaCLStatus ViennaCLCUDASsell_bicgstab(ViennaCLBackend backend, const ViennaCLCUDASsell *A, const float *b, float *x, ViennaCLB200SolverTagS *tag);
ViennaCLStatus ViennaCLCUDAScsr_gmres(ViennaCLBackend backend, const ViennaCLCUDAScsr *A, const float *b, float *x, ViennaCLB200SolverTagS *tag);
ViennaCLStatus ViennaCLCUDASsell_gmres(ViennaCLBackend backend, const ViennaCLCUDASsell *A, const float *b, float *x, ViennaCLB200SolverTagS *tag);
/* ell_matrix / hyb_matrix overloads of solve() (cg.hpp:204-254 and siblings) */
ViennaCLStatus ViennaCLCUDASell_cg(ViennaCLBackend backend, const ViennaCLCUDASell *A, const float *b, float *x, ViennaCLB200SolverTagS *tag);
ViennaCLStatus ViennaCLCUDAShyb_cg(ViennaCLBackend backend, const ViennaCLCUDAShyb *A, const float *b, float *x, ViennaCLB200SolverTagS *tag);
ViennaCLStatus ViennaCLCUDASell_bicgstab(ViennaCLBackend backend, const ViennaCLCUDASell *A, const float *b, float *x, ViennaCLB200SolverTagS *tag);
ViennaCLStatus ViennaCLCUDAShyb_bicgstab(ViennaCLBackend backend, const ViennaCLCUDAShyb *A, const float *b, float *x, ViennaCLB200SolverTagS *tag);
ViennaCLStatus ViennaCLCUDASell_gmres(ViennaCLBackend backend, const ViennaCLCUDASell *A, const float *b, float *x, ViennaCLB200SolverTagS *tag);
ViennaCLStatus ViennaCLCUDAShyb_gmres(ViennaCLBackend backend, const ViennaCLCUDAShyb *A, const float *b, float *x, ViennaCLB200SolverTagS *tag);

#ifdef __cplusplus
}
#endif
#endif /* VCL_B200_FLOAT_H */
