/* oracle/vcl_oracle.h -- TEST INFRASTRUCTURE ONLY (see vcl_oracle.c). */
#ifndef VCL_ORACLE_H
#define VCL_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned int vclo_u32;

/* ---- synthetic matrices: natural ordering, x fastest, Dirichlet neighbours dropped, ascending columns ---- */
long long vclo_gen_stencil2d(int nx, int ny, double cx, double cy, vclo_u32 *rp, vclo_u32 *ci, double *v);
long long vclo_gen_stencil3d(int nx, int ny, int nz, double cx, double cy, double cz, vclo_u32 *rp, vclo_u32 *ci, double *v);
void      vclo_fill_uniform(double *x, long long n, unsigned long long seed, double lo, double hi);

/* ---- SpMV ---- */
void vclo_csr_spmv(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const double *v,
                   const double *x, int offx, int incx, double alpha,
                   double *y, int offy, int incy, double beta);
long long vclo_sell_padded_nnz(int rows, const vclo_u32 *rp, int C);
void vclo_sell_build(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const double *v, int C,
                     vclo_u32 *cols_per_block, vclo_u32 *block_start, vclo_u32 *col_idx, double *elements);
void vclo_sell_spmv(int rows, int C, const vclo_u32 *cols_per_block, const vclo_u32 *block_start,
                    const vclo_u32 *col_idx, const double *elements,
                    const double *x, int offx, int incx, double alpha,
                    double *y, int offy, int incy, double beta);
/* ELL / HYB (AlignmentV = 1): layouts of ell_matrix.hpp:122-166 / hyb_matrix.hpp:127-214 */
int  vclo_ell_width(int rows, const vclo_u32 *rp);
void vclo_ell_build(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const double *v, int width, vclo_u32 *coords, double *elements);
void vclo_ell_spmv(int rows, int width, const vclo_u32 *coords, const double *elements,
                   const double *x, int offx, int incx, double alpha, double *y, int offy, int incy, double beta);
int  vclo_hyb_width(int rows, int cols, const vclo_u32 *rp, double threshold);
long long vclo_hyb_tail_nnz(int rows, const vclo_u32 *rp, int width);
void vclo_hyb_build(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const double *v, int width,
                    vclo_u32 *ell_coords, double *ell_elements, vclo_u32 *csr_rows, vclo_u32 *csr_cols, double *csr_elements);
void vclo_hyb_spmv(int rows, int width, const vclo_u32 *ell_coords, const double *ell_elements,
                   const vclo_u32 *csr_rows, const vclo_u32 *csr_cols, const double *csr_elements,
                   const double *x, int offx, int incx, double alpha, double *y, int offy, int incy, double beta);
void vclo_coo_spmv(int rows, long long nnz, const vclo_u32 *coords, const double *elements,
                   const double *x, double alpha, double *y, double beta);
void vclo_csr_diag(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const double *v, double *diag);

/* ---- BLAS-1 ---- */
double vclo_norm2(const double *x, long long n);
double vclo_inner_prod(const double *x, const double *y, long long n);

/* ---- solvers.  hist/hist_len optional (monitor estimates).  Return 0 on success. ---- */
int vclo_cg(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const double *v,
            const double *b, double *x, double tol, double abs_tol, int maxit,
            int *iters, double *err, double *hist, int hist_cap, int *hist_len);
int vclo_bicgstab(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const double *v,
                  const double *b, double *x, double tol, double abs_tol, int maxit,
                  int *iters, double *err, double *hist, int hist_cap, int *hist_len);
/* precond: 1 = Jacobi, 2 = identity */
int vclo_bicgstab_precond(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const double *v, int precond,
                          const double *b, double *x, double tol, double abs_tol, int maxit, int restart_every,
                          int *iters, double *err, double *hist, int hist_cap, int *hist_len);
int vclo_gmres(int rows, const vclo_u32 *rp, const vclo_u32 *ci, const double *v,
               const double *b, double *x, double tol, double abs_tol, int maxit, int krylov,
               int *iters, double *err, double *hist, int hist_cap, int *hist_len);

/* ---- multi-threaded timing helpers for bench.py's "port" baseline ---- */
int    vclo_max_threads(void);
void   vclo_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
