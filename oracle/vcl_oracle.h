/* oracle/vcl_oracle.h -- TEST INFRASTRUCTURE ONLY (see vcl_oracle.c). */
#ifndef VCL_ORACLE_H
#define VCL_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned int vclo_u32;
/* element type: the file is compiled twice, as it is (double, libvcl_oracle.so) and with -DVCLO_F32 (float,
 * libvcl_oracle_f32.so; the reference instantiates every type on the path for NumericT = float | double).
 * Tolerances, the reported error and the monitor history stay double (the reference's tags hold double, cg.hpp:48-87). */
#ifdef VCLO_F32
typedef float vreal;
#else
typedef double vreal;
#endif

/* ---- synthetic matrices: natural ordering, x fastest, Dirichlet neighbours dropped, ascending columns ---- */
long long vclo_gen_stencil2d(int nx, int ny, vreal cx, vreal cy, vclo_u32 *rp, vclo_u32 *ci, vreal *v);
long long vclo_gen_stencil3d(int nx, int ny, int nz, vreal cx, vreal cy, vreal cz, vclo_u32 *rp, vclo_u32 *ci, vreal *v);
void      vclo_fill_uniform(vreal *x, long long n, unsigned long long seed, vreal lo, vreal hi);

/* ---- SpMV ---- */
void vclo_csr_spmv(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const vreal *v,
                   const vreal *x, int offx, int incx, vreal alpha,
                   vreal *y, int offy, int incy, vreal beta);
long long vclo_sell_padded_nnz(int rows, const vclo_u32 *rp, int C);
void vclo_sell_build(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const vreal *v, int C,
                     vclo_u32 *cols_per_block, vclo_u32 *block_start, vclo_u32 *col_idx, vreal *elements);
void vclo_sell_spmv(int rows, int C, const vclo_u32 *cols_per_block, const vclo_u32 *block_start,
                    const vclo_u32 *col_idx, const vreal *elements,
                    const vreal *x, int offx, int incx, vreal alpha,
                    vreal *y, int offy, int incy, vreal beta);
/* ELL / HYB (AlignmentV = 1): layouts of ell_matrix.hpp:122-166 / hyb_matrix.hpp:127-214 */
int  vclo_ell_width(int rows, const vclo_u32 *rp);
void vclo_ell_build(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const vreal *v, int width, vclo_u32 *coords, vreal *elements);
void vclo_ell_spmv(int rows, int width, const vclo_u32 *coords, const vreal *elements,
                   const vreal *x, int offx, int incx, vreal alpha, vreal *y, int offy, int incy, vreal beta);
int  vclo_hyb_width(int rows, int cols, const vclo_u32 *rp, double threshold);
long long vclo_hyb_tail_nnz(int rows, const vclo_u32 *rp, int width);
void vclo_hyb_build(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const vreal *v, int width,
                    vclo_u32 *ell_coords, vreal *ell_elements, vclo_u32 *csr_rows, vclo_u32 *csr_cols, vreal *csr_elements);
void vclo_hyb_spmv(int rows, int width, const vclo_u32 *ell_coords, const vreal *ell_elements,
                   const vclo_u32 *csr_rows, const vclo_u32 *csr_cols, const vreal *csr_elements,
                   const vreal *x, int offx, int incx, vreal alpha, vreal *y, int offy, int incy, vreal beta);
void vclo_coo_spmv(int rows, long long nnz, const vclo_u32 *coords, const vreal *elements,
                   const vreal *x, vreal alpha, vreal *y, vreal beta);
void vclo_csr_diag(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const vreal *v, vreal *diag);

/* ---- BLAS-1 ---- */
vreal vclo_norm2(const vreal *x, long long n);
vreal vclo_inner_prod(const vreal *x, const vreal *y, long long n);

/* ---- solvers.  hist/hist_len optional (monitor estimates).  Return 0 on success. ---- */
int vclo_cg(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const vreal *v,
            const vreal *b, vreal *x, double tol, double abs_tol, int maxit,
            int *iters, double *err, double *hist, int hist_cap, int *hist_len);
int vclo_bicgstab(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const vreal *v,
                  const vreal *b, vreal *x, double tol, double abs_tol, int maxit,
                  int *iters, double *err, double *hist, int hist_cap, int *hist_len);
/* precond: 1 = Jacobi, 2 = identity */
int vclo_bicgstab_precond(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const vreal *v, int precond,
                          const vreal *b, vreal *x, double tol, double abs_tol, int maxit, int restart_every,
                          int *iters, double *err, double *hist, int hist_cap, int *hist_len);
int vclo_gmres(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const vreal *v,
               const vreal *b, vreal *x, double tol, double abs_tol, int maxit, int krylov,
               int *iters, double *err, double *hist, int hist_cap, int *hist_len);

/* ---- multi-threaded timing helpers for bench.py's "port" baseline ---- */
int    vclo_max_threads(void);
void   vclo_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
