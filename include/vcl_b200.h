/* vcl_b200.h -- C-ABI of libvcl_b200.so: the B200-native (sm_100a) backend for ViennaCL's sparse SpMV + Krylov hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one `case viennacl::CUDA_MEMORY:` arm of
 * the reference's dispatch layer; the replaced reference interface is cited as file:line (relative to the reference tree).
 * Style follows libviennacl (libviennacl/include/viennacl.hpp:44-113): C linkage, ViennaCLStatus return, ViennaCLInt sizes,
 * raw DEVICE pointers, precision letter in the name, opaque ViennaCLBackend first.
 *
 * Conventions
 *   - all vector / matrix pointers are device pointers on the backend's device unless the name says "Host";
 *   - indices are `unsigned int` exactly as in the reference (compressed_matrix.hpp:1190-1197);
 *   - calls are asynchronous on the backend's stream unless documented otherwise; there is no CPU fallback:
 *     every compute entry point fails with ViennaCLB200NoDevice when no sm_100 device is usable;
 *   - the caller owns every buffer passed in; the library never frees or retains user pointers
 *     (wrap semantics of compressed_matrix.hpp:740-781, vector.hpp:271-293).
 */
#ifndef VCL_B200_H
#define VCL_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int ViennaCLInt;

/* libviennacl/include/viennacl.hpp:98-102 -- extended with specific failure codes (all non-zero = failure). */
typedef enum
{
  ViennaCLSuccess = 0,
  ViennaCLGenericFailure,
  ViennaCLB200InvalidArgument,
  ViennaCLB200CudaError,
  ViennaCLB200NoDevice,
  ViennaCLB200OutOfMemory,
  ViennaCLB200NotInitialized,
  ViennaCLB200CommError
} ViennaCLStatus;

/* libviennacl/src/viennacl_private.hpp:38-62: "TODO: Add stream and/or device descriptors here" -- done here. */
struct ViennaCLBackend_impl;
typedef struct ViennaCLBackend_impl *ViennaCLBackend;

/* ---------------------------------------------------------------- backend ---------------------------------------------------------------- */
/* libviennacl/src/backend.cpp:24-46 */
ViennaCLStatus ViennaCLBackendCreate(ViennaCLBackend *backend);                       /* device = current, own non-blocking stream */
ViennaCLStatus ViennaCLBackendCreateOnDevice(ViennaCLBackend *backend, ViennaCLInt device, void *cuda_stream /* NULL: create one */);
ViennaCLStatus ViennaCLBackendDestroy(ViennaCLBackend *backend);
ViennaCLStatus ViennaCLBackendSynchronize(ViennaCLBackend backend);                   /* backend::finish(), backend/memory.hpp:54-62 */
ViennaCLStatus ViennaCLBackendGetStream(ViennaCLBackend backend, void **cuda_stream);
ViennaCLStatus ViennaCLBackendGetDevice(ViennaCLBackend backend, ViennaCLInt *device, ViennaCLInt *sm_count);
const char    *ViennaCLBackendLastError(ViennaCLBackend backend);                     /* text of the last failure on this handle */
const char    *ViennaCLB200Version(void);
/* Timing on the backend's stream (CUDA events): Begin, enqueue work, End returns elapsed milliseconds (synchronises). */
ViennaCLStatus ViennaCLBackendTimerBegin(ViennaCLBackend backend);
ViennaCLStatus ViennaCLBackendTimerEnd(ViennaCLBackend backend, double *milliseconds);
/* Writes a scratch buffer larger than L2 so that the next timed call starts cold. */
ViennaCLStatus ViennaCLBackendFlushL2(ViennaCLBackend backend);
/* Per-handle tuning knobs (all have working defaults):
 *   "persistent_rows"  largest system (rows) solved by the persistent cooperative kernels (whole CG / BiCGStab iterations and
 *                      GMRES cycles inside one launch); larger systems take the multi-kernel drivers.  -1: built-in default
 *                      (10M rows; env VCL_B200_PERSISTENT_ROWS at handle creation), 0: always multi-kernel.
 *   "l2_resident"      persistent kernels keep a matrix that fits L2 resident there (evict-last) instead of streaming it
 *                      (evict-first): 0 never (default: measured slower on B200), 1 evict-last, 2 normal policy.
 *   "persistent_cg_form"  0 (default): chosen by size; 1: one-pass persistent CG -- one grid barrier per iteration, the product
 *                      recomputes the updated search direction on the fly (best up to ~600k rows); 2: the two-phase form (update,
 *                      barrier, product, barrier); 3: one-pass compiled for 3 CTAs per SM. */
ViennaCLStatus ViennaCLBackendSetOption(ViennaCLBackend backend, const char *name, long long value);
/* Counts kernels launched by this library on this handle since creation (bench.py's gpu_launches). */
ViennaCLStatus ViennaCLBackendLaunchCount(ViennaCLBackend backend, long long *launches);

/* Multi-GPU (one process per GPU).  The unique id is created on rank 0 and broadcast by the host program
 * (torch.distributed / MPI / files); not present in the reference (doc/manual/multi-device.dox:9). */
#define VCL_B200_COMM_ID_BYTES 128
ViennaCLStatus ViennaCLBackendCommGetUniqueId(ViennaCLBackend backend, void *id_bytes /* VCL_B200_COMM_ID_BYTES */);
ViennaCLStatus ViennaCLBackendCommInit(ViennaCLBackend backend, const void *id_bytes, ViennaCLInt rank, ViennaCLInt world_size);
ViennaCLStatus ViennaCLBackendCommDestroy(ViennaCLBackend backend);
/* Polls the communicator for asynchronous errors (ncclCommGetAsyncError); the NCCL transport does so itself once per solver batch. */
ViennaCLStatus ViennaCLBackendCommCheck(ViennaCLBackend backend);

/* ---------------------------------------------------------------- memory ----------------------------------------------------------------- */
/* backend/cuda.hpp:103-200 (memory_create / memory_copy / memory_write / memory_read) */
ViennaCLStatus ViennaCLCUDAMemAlloc(ViennaCLBackend backend, void **ptr, size_t bytes);
ViennaCLStatus ViennaCLCUDAMemFree(ViennaCLBackend backend, void *ptr);
ViennaCLStatus ViennaCLCUDAMemWrite(ViennaCLBackend backend, void *dst_dev, size_t dst_offset, const void *src_host, size_t bytes, ViennaCLInt async);
ViennaCLStatus ViennaCLCUDAMemRead(ViennaCLBackend backend, const void *src_dev, size_t src_offset, void *dst_host, size_t bytes, ViennaCLInt async);
ViennaCLStatus ViennaCLCUDAMemCopy(ViennaCLBackend backend, const void *src_dev, size_t src_offset, void *dst_dev, size_t dst_offset, size_t bytes);
ViennaCLStatus ViennaCLCUDAMemSet(ViennaCLBackend backend, void *dst_dev, ViennaCLInt byte_value, size_t bytes);
ViennaCLStatus ViennaCLHostAllocPinned(ViennaCLBackend backend, void **ptr, size_t bytes);
ViennaCLStatus ViennaCLHostFreePinned(ViennaCLBackend backend, void *ptr);

/* ---------------------------------------------------------------- BLAS-1 subset ----------------------------------------------------------- */
/* linalg/cuda/vector_operations.hpp:77 (av), :179 (avbv), :483 (avbv_v), :782 (vector_assign), :870 (element_op /),
 * :1273-1579 (inner_prod + inner_prod_cpu), :2018-2448 (norm_2 + norm_2_cpu). */
ViennaCLStatus ViennaCLCUDADav(ViennaCLBackend backend, ViennaCLInt n, double *x, ViennaCLInt offx, ViennaCLInt incx,
                               const double *y, ViennaCLInt offy, ViennaCLInt incy, double alpha);                     /* x = alpha*y */
ViennaCLStatus ViennaCLCUDADavbv(ViennaCLBackend backend, ViennaCLInt n, double *x, ViennaCLInt offx, ViennaCLInt incx,
                                 const double *y, ViennaCLInt offy, ViennaCLInt incy, double alpha,
                                 const double *z, ViennaCLInt offz, ViennaCLInt incz, double beta);                    /* x = alpha*y + beta*z */
ViennaCLStatus ViennaCLCUDADavbv_v(ViennaCLBackend backend, ViennaCLInt n, double *x, ViennaCLInt offx, ViennaCLInt incx,
                                   const double *y, ViennaCLInt offy, ViennaCLInt incy, double alpha,
                                   const double *z, ViennaCLInt offz, ViennaCLInt incz, double beta);                  /* x += alpha*y + beta*z */
ViennaCLStatus ViennaCLCUDADassign(ViennaCLBackend backend, ViennaCLInt n, double *x, ViennaCLInt offx, ViennaCLInt incx, double value);
ViennaCLStatus ViennaCLCUDADelement_div(ViennaCLBackend backend, ViennaCLInt n, double *x, ViennaCLInt offx, ViennaCLInt incx,
                                        const double *y, ViennaCLInt offy, ViennaCLInt incy,
                                        const double *z, ViennaCLInt offz, ViennaCLInt incz);                          /* x = y ./ z */
ViennaCLStatus ViennaCLCUDADdot(ViennaCLBackend backend, ViennaCLInt n, double *result_host,
                                const double *x, ViennaCLInt offx, ViennaCLInt incx,
                                const double *y, ViennaCLInt offy, ViennaCLInt incy);                                  /* synchronous */
ViennaCLStatus ViennaCLCUDADnrm2(ViennaCLBackend backend, ViennaCLInt n, double *result_host,
                                 const double *x, ViennaCLInt offx, ViennaCLInt incx);                                 /* synchronous */

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
ViennaCLStatus ViennaCLCUDAcsr_row_blocks(ViennaCLBackend backend, ViennaCLInt rows, const unsigned int *row_ptr,
                                          unsigned int *row_blocks, ViennaCLInt *num_blocks);

/* y[offy + i*incy] = alpha * (A x)_i + (beta != 0 ? beta * y[...] : 0);  x read at offx + col*incx.
 * linalg/sparse_matrix_operations.hpp:90-121 -> cuda/sparse_matrix_operations.hpp:262-396 (kernels :137-249).
 * row_blocks: a plan from ViennaCLCUDAcsr_row_blocks, or ANY partition of the rows into consecutive blocks (e.g. the reference's own
 * handle3() plan) -- a plan this library did not make is checked once per plan address (cached; rewriting or freeing the plan's
 * memory through this API drops the verdict) and, when a block breaks the limits above, the product runs through the plan-free
 * kernel (one thread per row, no staging) -- as it does for row_blocks == NULL or arrays that are not 16-byte aligned. */
ViennaCLStatus ViennaCLCUDADcsrmv(ViennaCLBackend backend, ViennaCLInt rows, ViennaCLInt cols, ViennaCLInt nnz,
                                  const unsigned int *row_ptr, const unsigned int *col_idx, const double *values,
                                  const unsigned int *row_blocks, ViennaCLInt num_blocks,
                                  const double *x, ViennaCLInt offx, ViennaCLInt incx, double alpha,
                                  double *y, ViennaCLInt offy, ViennaCLInt incy, double beta);

/* SELL-C-sigma (sigma = 1): cuda/sparse_matrix_operations.hpp:2196-2289; layout sliced_ell_matrix.hpp:134-214. */
ViennaCLStatus ViennaCLCUDADsellmv(ViennaCLBackend backend, ViennaCLInt rows, ViennaCLInt cols, ViennaCLInt rows_per_block,
                                   const unsigned int *columns_per_block, const unsigned int *col_idx,
                                   const unsigned int *block_start, const double *values,
                                   const double *x, ViennaCLInt offx, ViennaCLInt incx, double alpha,
                                   double *y, ViennaCLInt offy, ViennaCLInt incy, double beta);

/* Device-side CSR -> SELL-C conversion with the exact array layout of sliced_ell_matrix.hpp:140-214.
 * Call 1: columns_per_block/block_start sized ceil(rows/C) are filled and *padded_nnz returned (col_idx/values NULL).
 * Call 2: col_idx/values sized *padded_nnz are filled. */
ViennaCLStatus ViennaCLCUDADcsr2sell(ViennaCLBackend backend, ViennaCLInt rows, ViennaCLInt rows_per_block,
                                     const unsigned int *row_ptr, const unsigned int *csr_col, const double *csr_val,
                                     unsigned int *columns_per_block, unsigned int *block_start, long long *padded_nnz,
                                     unsigned int *col_idx, double *values);
/* SELL-C-sigma (not in the reference, whose sigma is fixed at 1): inside every window of `sigma` consecutive rows
 * (sigma a multiple of rows_per_block, <= 4096) the rows are ordered by decreasing length -- stable, ties keep their order --
 * before they are cut into slices, which removes most of the padding of irregular matrices.  row_perm (out, ceil(rows/C)*C
 * entries) maps storage rows to matrix rows.  Same two-call protocol as csr2sell: the first call (col_idx == NULL) fills
 * row_perm, columns_per_block, block_start and *padded_nnz, the second one scatters the entries.  Products and solvers take
 * the matrix through the ViennaCLCUDADsell struct (row_perm set); per-row arithmetic and results equal those of sigma = 1. */
ViennaCLStatus ViennaCLCUDADcsr2sell_sigma(ViennaCLBackend backend, ViennaCLInt rows, ViennaCLInt rows_per_block, ViennaCLInt sigma,
                                           const unsigned int *row_ptr, const unsigned int *csr_col, const double *csr_val,
                                           unsigned int *row_perm, unsigned int *columns_per_block, unsigned int *block_start,
                                           long long *padded_nnz, unsigned int *col_idx, double *values);

/* ELL (ell_matrix.hpp:36-119) and HYB (hyb_matrix.hpp:36-126), AlignmentV = 1 layouts:
 * ELL entry j of row r at j*internal_rows + r (coords / elements hold internal_rows*maxnnz entries, padding value 0, column 0);
 * HYB = ELL part of width ell.maxnnz + CSR tail (csr_rows[rows+1], csr_cols, csr_elements).
 * Products: cuda/sparse_matrix_operations.hpp:1747-1838 (ELL), :2298-2400 (HYB); zero-valued ELL slots never touch x. */
typedef struct
{
  ViennaCLInt rows, cols, internal_rows, maxnnz;
  const unsigned int *coords;
  const double *elements;
} ViennaCLCUDADell;

typedef struct
{
  ViennaCLCUDADell ell;
  const unsigned int *csr_rows, *csr_cols;
  const double *csr_elements;
  ViennaCLInt csr_nnz;
} ViennaCLCUDADhyb;

ViennaCLStatus ViennaCLCUDADellmv(ViennaCLBackend backend, const ViennaCLCUDADell *A,
                                  const double *x, ViennaCLInt offx, ViennaCLInt incx, double alpha,
                                  double *y, ViennaCLInt offy, ViennaCLInt incy, double beta);
ViennaCLStatus ViennaCLCUDADhybmv(ViennaCLBackend backend, const ViennaCLCUDADhyb *A,
                                  const double *x, ViennaCLInt offx, ViennaCLInt incx, double alpha,
                                  double *y, ViennaCLInt offy, ViennaCLInt incy, double beta);
/* Device-side CSR -> ELL (layout of ell_matrix.hpp:122-166).  Call 1 (coords == NULL): *maxnnz = longest row.
 * Call 2: coords / elements sized rows * (*maxnnz) are filled (internal_rows = rows). */
ViennaCLStatus ViennaCLCUDADcsr2ell(ViennaCLBackend backend, ViennaCLInt rows, const unsigned int *row_ptr,
                                    const unsigned int *csr_col, const double *csr_val, ViennaCLInt *maxnnz,
                                    unsigned int *coords, double *elements);
/* Device-side CSR -> HYB (layout and width rule of hyb_matrix.hpp:127-214: the smallest width that covers at least
 * `csr_threshold` (reference default 0.8) of the rows).  Call 1 (ell_coords == NULL): *ell_width and *csr_nnz (>= 1: the
 * reference stores one dummy entry when the tail is empty).  Call 2 fills all five arrays. */
ViennaCLStatus ViennaCLCUDADcsr2hyb(ViennaCLBackend backend, ViennaCLInt rows, ViennaCLInt cols, const unsigned int *row_ptr,
                                    const unsigned int *csr_col, const double *csr_val, double csr_threshold,
                                    ViennaCLInt *ell_width, ViennaCLInt *csr_nnz,
                                    unsigned int *ell_coords, double *ell_elements,
                                    unsigned int *csr_rows, unsigned int *csr_cols, double *csr_elements);

/* COO (coordinate_matrix.hpp:47-102): coords = (row, col) pairs, entries sorted by row as the reference's copy() produces
 * them.  SpMV and the solvers run on a CSR index of the same entries, built once on the device:
 * row_ptr[rows+1] and col_idx[nnz] are filled, the value array is shared with the COO matrix (no copy).
 * Fails with ViennaCLB200InvalidArgument when the entries are not sorted by row. */
ViennaCLStatus ViennaCLCUDAcoo2csr(ViennaCLBackend backend, ViennaCLInt rows, ViennaCLInt nnz, const unsigned int *coords,
                                   unsigned int *row_ptr, unsigned int *col_idx);
/* coordinate_matrix product with the reference's arithmetic (host_based/sparse_matrix_operations.hpp:1222-1247):
 * y <- beta*y (or 0), then y[row] += (alpha*a) * x[col] entry by entry; (row_ptr, col_idx) from ViennaCLCUDAcoo2csr. */
ViennaCLStatus ViennaCLCUDADcoomv(ViennaCLBackend backend, ViennaCLInt rows, ViennaCLInt cols, ViennaCLInt nnz,
                                  const unsigned int *row_ptr, const unsigned int *col_idx, const double *elements,
                                  const unsigned int *row_blocks, ViennaCLInt num_blocks,
                                  const double *x, ViennaCLInt offx, ViennaCLInt incx, double alpha,
                                  double *y, ViennaCLInt offy, ViennaCLInt incy, double beta);

/* detail::row_info: linalg/sparse_matrix_operations.hpp:48-74 -> cuda/sparse_matrix_operations.hpp:53-119.
 * option: 0 inf-norm, 1 1-norm, 2 2-norm, 3 diagonal (forwards.h row_info_types order). */
ViennaCLStatus ViennaCLCUDADcsr_row_info(ViennaCLBackend backend, ViennaCLInt rows,
                                         const unsigned int *row_ptr, const unsigned int *col_idx, const double *values,
                                         double *result, ViennaCLInt option);

/* ---------------------------------------------------------------- matrix / vector generators ---------------------------------------------- */
/* tools/matrix_generation.hpp:47-88 generalised (DESIGN.md "synthetic inputs"): 5-/7-point stencil with first-order upwind
 * convection c; nz == 1 selects the 2-D 5-point stencil.  row_ptr[rows+1], col_idx[nnz], values[nnz] are device buffers;
 * pass NULLs to get the counts only. */
ViennaCLStatus ViennaCLCUDADgenerate_stencil(ViennaCLBackend backend, ViennaCLInt nx, ViennaCLInt ny, ViennaCLInt nz,
                                             double cx, double cy, double cz,
                                             unsigned int *row_ptr, unsigned int *col_idx, double *values,
                                             long long *rows, long long *nnz);
/* Same matrix restricted to rows [row_begin, row_end) (global column indices, local row_ptr starting at 0). */
ViennaCLStatus ViennaCLCUDADgenerate_stencil_rows(ViennaCLBackend backend, ViennaCLInt nx, ViennaCLInt ny, ViennaCLInt nz,
                                                  double cx, double cy, double cz, long long row_begin, long long row_end,
                                                  unsigned int *row_ptr, unsigned int *col_idx, double *values, long long *nnz);
ViennaCLStatus ViennaCLCUDADfill_uniform(ViennaCLBackend backend, long long n, double *x, unsigned long long seed,
                                         long long index_offset, double lo, double hi);

/* ---------------------------------------------------------------- fused solver steps ------------------------------------------------------ */
/* One-to-one with linalg/iterative_operations.hpp (argument lists :59-65, :97-100, :134-139, :171-176, :208-214, :248-255,
 * :286-291, :321-329, :356-362, :393-396).  `buf` is the reference's inner_prod_buffer: chunks of `chunk` entries; each
 * routine writes the FULLY REDUCED value into element 0 of its chunk(s) and leaves the rest untouched (the reference's
 * host backend does the same, host_based/iterative_operations.hpp:100-102; the drivers sum whole chunks). */
typedef struct
{
  ViennaCLInt rows, cols, nnz;
  const unsigned int *row_ptr, *col_idx;
  const double *values;
  const unsigned int *row_blocks;   /* may be NULL */
  ViennaCLInt num_blocks;
} ViennaCLCUDADcsr;

typedef struct
{
  ViennaCLInt rows, cols, rows_per_block;
  const unsigned int *columns_per_block, *col_idx, *block_start;
  const double *values;
  const unsigned int *row_perm;   /* NULL: sigma = 1, the reference's layout (sliced_ell_matrix.hpp:43).  Otherwise SELL-C-sigma:
                                     storage row i holds matrix row row_perm[i] (0xFFFFFFFF: padding row), see csr2sell_sigma */
} ViennaCLCUDADsell;
/* y = alpha*A*x + beta*y for a SELL matrix passed as the struct (needed for SELL-C-sigma: row_perm); ViennaCLCUDADsellmv is the
 * flat-argument form of the reference's kernel signature. */
ViennaCLStatus ViennaCLCUDADsellmv_struct(ViennaCLBackend backend, const ViennaCLCUDADsell *A,
                                          const double *x, ViennaCLInt offx, ViennaCLInt incx, double alpha,
                                          double *y, ViennaCLInt offy, ViennaCLInt incy, double beta);

ViennaCLStatus ViennaCLCUDADpipelined_cg_vector_update(ViennaCLBackend backend, ViennaCLInt n, double *result, double alpha,
                                                       double *p, double *r, const double *Ap, double beta,
                                                       double *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDADpipelined_cg_prod_csr(ViennaCLBackend backend, const ViennaCLCUDADcsr *A, const double *p, double *Ap,
                                                  double *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDADpipelined_cg_prod_sell(ViennaCLBackend backend, const ViennaCLCUDADsell *A, const double *p, double *Ap,
                                                   double *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDADpipelined_bicgstab_update_s(ViennaCLBackend backend, ViennaCLInt n, double *s, const double *r, const double *Ap,
                                                        double *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset);
ViennaCLStatus ViennaCLCUDADpipelined_bicgstab_vector_update(ViennaCLBackend backend, ViennaCLInt n, double *result, double alpha, double *p,
                                                             double omega, const double *s, double *residual, const double *As,
                                                             double beta, const double *Ap, const double *r0star,
                                                             double *buf, ViennaCLInt chunk);
ViennaCLStatus ViennaCLCUDADpipelined_bicgstab_prod_csr(ViennaCLBackend backend, const ViennaCLCUDADcsr *A, const double *p, double *Ap,
                                                        const double *r0star, double *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset);
ViennaCLStatus ViennaCLCUDADpipelined_bicgstab_prod_sell(ViennaCLBackend backend, const ViennaCLCUDADsell *A, const double *p, double *Ap,
                                                         const double *r0star, double *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset);
ViennaCLStatus ViennaCLCUDADpipelined_gmres_normalize_vk(ViennaCLBackend backend, ViennaCLInt n, double *v_k, const double *residual,
                                                         double *R, ViennaCLInt offset_in_R, const double *buf,
                                                         double *r_dot_vk, ViennaCLInt chunk, ViennaCLInt chunk_offset);
ViennaCLStatus ViennaCLCUDADpipelined_gmres_gram_schmidt_stage1(ViennaCLBackend backend, const double *basis, ViennaCLInt n,
                                                                ViennaCLInt internal_n, ViennaCLInt k, double *vi_in_vk, ViennaCLInt chunk);
ViennaCLStatus ViennaCLCUDADpipelined_gmres_gram_schmidt_stage2(ViennaCLBackend backend, double *basis, ViennaCLInt n,
                                                                ViennaCLInt internal_n, ViennaCLInt k, const double *vi_in_vk,
                                                                double *R, ViennaCLInt krylov_dim, double *buf, ViennaCLInt chunk);
ViennaCLStatus ViennaCLCUDADpipelined_gmres_update_result(ViennaCLBackend backend, ViennaCLInt n, double *result, const double *residual,
                                                          const double *basis, ViennaCLInt internal_n, const double *coefficients, ViennaCLInt k);
ViennaCLStatus ViennaCLCUDADpipelined_gmres_prod_csr(ViennaCLBackend backend, const ViennaCLCUDADcsr *A, const double *p, double *Ap,
                                                     double *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDADpipelined_gmres_prod_sell(ViennaCLBackend backend, const ViennaCLCUDADsell *A, const double *p, double *Ap,
                                                      double *buf, ViennaCLInt buf_size);
/* the same fused products for ell_matrix / hyb_matrix (cuda/iterative_operations.hpp:330-727, :1138-1593) */
ViennaCLStatus ViennaCLCUDADpipelined_cg_prod_ell(ViennaCLBackend backend, const ViennaCLCUDADell *A, const double *p, double *Ap,
                                                  double *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDADpipelined_cg_prod_hyb(ViennaCLBackend backend, const ViennaCLCUDADhyb *A, const double *p, double *Ap,
                                                  double *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDADpipelined_bicgstab_prod_ell(ViennaCLBackend backend, const ViennaCLCUDADell *A, const double *p, double *Ap,
                                                        const double *r0star, double *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset);
ViennaCLStatus ViennaCLCUDADpipelined_bicgstab_prod_hyb(ViennaCLBackend backend, const ViennaCLCUDADhyb *A, const double *p, double *Ap,
                                                        const double *r0star, double *buf, ViennaCLInt chunk, ViennaCLInt chunk_offset);
ViennaCLStatus ViennaCLCUDADpipelined_gmres_prod_ell(ViennaCLBackend backend, const ViennaCLCUDADell *A, const double *p, double *Ap,
                                                     double *buf, ViennaCLInt buf_size);
ViennaCLStatus ViennaCLCUDADpipelined_gmres_prod_hyb(ViennaCLBackend backend, const ViennaCLCUDADhyb *A, const double *p, double *Ap,
                                                     double *buf, ViennaCLInt buf_size);

/* ---------------------------------------------------------------- whole solves ------------------------------------------------------------ */
/* The loop lives next to the kernels (device-resident scalars, no per-iteration host round trip); the C++ `solve()` keeps its
 * signature and calls these.  Semantics (tolerances, iteration counting, error estimate, quirks) are those of
 * linalg/cg.hpp:128-187, bicgstab.hpp:97-215 / :398-489 and gmres.hpp:181-367.
 * monitor (optional) has the reference's contract: called with the device pointer of the current iterate and the relative
 * residual estimate, once per iteration (GMRES: once per restart); returning non-zero stops the solver. */
typedef ViennaCLInt (*ViennaCLMonitorD)(const double *x_dev, double rel_residual_estimate, void *user);

/* Diagonal preconditioners that the drivers fold into their kernels (CSR matrices): Jacobi (jacobi_precond.hpp:103-130,
 * divide by diag(A)) and row scaling (row_scaling.hpp:150-190, divide by the inf-/1-/2-norm of the row). */
typedef enum { ViennaCLB200PrecondNone = 0, ViennaCLB200PrecondJacobi = 1, ViennaCLB200PrecondRowScalingInf = 2,
               ViennaCLB200PrecondRowScaling1 = 3, ViennaCLB200PrecondRowScaling2 = 4 } ViennaCLB200Precond;

typedef struct
{
  double tolerance;                 /* relative */
  double abs_tolerance;
  ViennaCLInt max_iterations;
  ViennaCLInt krylov_dim;           /* GMRES only */
  ViennaCLInt max_iterations_before_restart; /* BiCGStab (preconditioned path) only */
  ViennaCLB200Precond precond;
  ViennaCLMonitorD monitor;
  void *monitor_user;
  /* results */
  ViennaCLInt iters;
  double error;
} ViennaCLB200SolverTag;

ViennaCLStatus ViennaCLCUDADcsr_cg(ViennaCLBackend backend, const ViennaCLCUDADcsr *A, const double *b, double *x, ViennaCLB200SolverTag *tag);
ViennaCLStatus ViennaCLCUDADsell_cg(ViennaCLBackend backend, const ViennaCLCUDADsell *A, const double *b, double *x, ViennaCLB200SolverTag *tag);
ViennaCLStatus ViennaCLCUDADcsr_bicgstab(ViennaCLBackend backend, const ViennaCLCUDADcsr *A, const double *b, double *x, ViennaCLB200SolverTag *tag);
ViennaCLStatus ViennaCLCUDADsell_bicgstab(ViennaCLBackend backend, const ViennaCLCUDADsell *A, const double *b, double *x, ViennaCLB200SolverTag *tag);
ViennaCLStatus ViennaCLCUDADcsr_gmres(ViennaCLBackend backend, const ViennaCLCUDADcsr *A, const double *b, double *x, ViennaCLB200SolverTag *tag);
ViennaCLStatus ViennaCLCUDADsell_gmres(ViennaCLBackend backend, const ViennaCLCUDADsell *A, const double *b, double *x, ViennaCLB200SolverTag *tag);
/* ell_matrix / hyb_matrix overloads of solve() (cg.hpp:204-254 and siblings) */
ViennaCLStatus ViennaCLCUDADell_cg(ViennaCLBackend backend, const ViennaCLCUDADell *A, const double *b, double *x, ViennaCLB200SolverTag *tag);
ViennaCLStatus ViennaCLCUDADhyb_cg(ViennaCLBackend backend, const ViennaCLCUDADhyb *A, const double *b, double *x, ViennaCLB200SolverTag *tag);
ViennaCLStatus ViennaCLCUDADell_bicgstab(ViennaCLBackend backend, const ViennaCLCUDADell *A, const double *b, double *x, ViennaCLB200SolverTag *tag);
ViennaCLStatus ViennaCLCUDADhyb_bicgstab(ViennaCLBackend backend, const ViennaCLCUDADhyb *A, const double *b, double *x, ViennaCLB200SolverTag *tag);
ViennaCLStatus ViennaCLCUDADell_gmres(ViennaCLBackend backend, const ViennaCLCUDADell *A, const double *b, double *x, ViennaCLB200SolverTag *tag);
ViennaCLStatus ViennaCLCUDADhyb_gmres(ViennaCLBackend backend, const ViennaCLCUDADhyb *A, const double *b, double *x, ViennaCLB200SolverTag *tag);

/* ---------------------------------------------------------------- mixed precision (double outside, float inside) ------------------------- */
/* linalg/mixed_precision_cg.hpp:95-186: CG on a compressed_matrix<double> whose inner iterations run on a float copy of the
 * matrix (8 instead of 12 bytes per entry); whenever the float residual has dropped by `inner_tolerance` (a ratio of SQUARED
 * norms, :160) the iterate is folded into the double result, the residual is recomputed in double (:165-166) and the float
 * iteration restarts from it.  tag: tolerance, max_iterations (total inner iterations), iters/error as in the reference
 * (:151, :182).  Here the inner iterations are the fused single-precision pipelined CG (ViennaCLCUDAScsr_cg).
 * values_float: the float copy of A's values if the caller keeps one (compressed_matrix<float> sharing A's index arrays),
 * or NULL: converted into backend workspace on every call. */
ViennaCLStatus ViennaCLCUDAconvert_DtoS(ViennaCLBackend backend, long long n, const double *x, float *y);        /* y = (float) x */
ViennaCLStatus ViennaCLCUDAconvert_StoD(ViennaCLBackend backend, long long n, const float *x, double *y);        /* y = (double) x */
ViennaCLStatus ViennaCLCUDADcsr_mixed_precision_cg(ViennaCLBackend backend, const ViennaCLCUDADcsr *A, const float *values_float,
                                                   const double *b, double *x, float inner_tolerance, ViennaCLB200SolverTag *tag);

/* ---------------------------------------------------------------- row-partitioned (multi-GPU) solvers ------------------------------------- */
/* New (no reference counterpart).  Rank g owns a contiguous block of rows of a square matrix; `A_local` holds those rows with
 * GLOBAL column indices.  DistCreate analyses the halo (columns outside the owned range, assumed to belong to the two
 * neighbouring ranks only... general owners are supported through an all-to-all send list), remaps the columns to
 * [owned | halo] and splits rows into interior / boundary.  b_local / x_local are the owned slices. */
struct ViennaCLB200DistCsr_impl;
typedef struct ViennaCLB200DistCsr_impl *ViennaCLB200DistCsr;
ViennaCLStatus ViennaCLCUDADdist_csr_create(ViennaCLBackend backend, long long global_rows, long long row_begin, long long row_end,
                                            ViennaCLInt local_nnz, const unsigned int *row_ptr, const unsigned int *col_idx_global,
                                            const double *values, ViennaCLB200DistCsr *out);
ViennaCLStatus ViennaCLCUDADdist_csr_destroy(ViennaCLBackend backend, ViennaCLB200DistCsr *A);
ViennaCLStatus ViennaCLCUDADdist_csrmv(ViennaCLBackend backend, ViennaCLB200DistCsr A, const double *x_local, double *y_local);
/* Storage format of the slab used by the products and solver steps: 0 = CSR (default), 1 = SELL-C with sigma = 1 (the layout of
 * sliced_ell_matrix.hpp:134-214, built on the device; rows_per_block <= 0: 32).  With SELL the per-row arithmetic is the reference's
 * SELL arithmetic: results equal the single-domain SELL product bit for bit. */
ViennaCLStatus ViennaCLCUDADdist_csr_set_format(ViennaCLBackend backend, ViennaCLB200DistCsr A, ViennaCLInt format, ViennaCLInt rows_per_block);
/* Which transport the object uses: 1 = peer memory (CUDA IPC windows over NVLink; halo pushes and the reduction are done
 * by the solver's own kernels), 0 = NCCL send/recv + allreduce.  Also returns the halo size and the interior/boundary split. */
ViennaCLStatus ViennaCLCUDADdist_csr_info(ViennaCLBackend backend, ViennaCLB200DistCsr A, ViennaCLInt *peer_memory,
                                          ViennaCLInt *halo_entries, ViennaCLInt *interior_blocks, ViennaCLInt *boundary_blocks);
/* cg.hpp:128-187 over slabs; tag->precond: none, Jacobi or row scaling (diagonal preconditioners, single-reduction PCG as in the
 * single-GPU library). */
ViennaCLStatus ViennaCLCUDADdist_csr_cg(ViennaCLBackend backend, ViennaCLB200DistCsr A, const double *b_local, double *x_local,
                                        ViennaCLB200SolverTag *tag);
/* bicgstab.hpp:97-215 (pipelined, no preconditioner) over slabs: two halo exchanges and two all-reduces per iteration, both inside
 * the product kernels on the peer-memory transport. */
ViennaCLStatus ViennaCLCUDADdist_csr_bicgstab(ViennaCLBackend backend, ViennaCLB200DistCsr A, const double *b_local, double *x_local,
                                              ViennaCLB200SolverTag *tag);
/* gmres.hpp:181-367 (pipelined GMRES(m), no preconditioner) over slabs: the basis is partitioned like the vectors; the k Gram-Schmidt
 * dots of an inner iteration are all-reduced together. */
ViennaCLStatus ViennaCLCUDADdist_csr_gmres(ViennaCLBackend backend, ViennaCLB200DistCsr A, const double *b_local, double *x_local,
                                           ViennaCLB200SolverTag *tag);

#ifdef __cplusplus
}
#endif
#endif /* VCL_B200_H */
