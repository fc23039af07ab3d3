/* oracle/vcl_oracle_mixed.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see vcl_oracle.c).
 *
 * Plain-C restatement of viennacl/linalg/mixed_precision_cg.hpp:95-186: CG on a double system whose inner iterations
 * (classical CG: SpMV, <Ap,p>, x += alpha p, r -= alpha Ap, <r,r>, p = r + beta p) run in float on a float copy of the
 * matrix; when the float residual has dropped by inner_tol (ratio of SQUARED norms, :160) or the budget is exhausted, the
 * float iterate is folded into the double result, the residual is recomputed in double and the float iteration restarts.
 * Linked into libvcl_oracle.so (the double build) only.  Pinned against the reference build by tests/test_oracle.py. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned int u32;

static void spmv_f(int rows, const u32 *rp, const u32 *ci, const float *v, const float *x, float *y)
{
  for (int r = 0; r < rows; ++r)
  {
    float dot = 0;
    for (u32 k = rp[r]; k < rp[r + 1]; ++k) dot += v[k] * x[ci[k]];
    y[r] = dot;
  }
}

static void spmv_d(int rows, const u32 *rp, const u32 *ci, const double *v, const double *x, double *y)
{
  for (int r = 0; r < rows; ++r)
  {
    double dot = 0;
    for (u32 k = rp[r]; k < rp[r + 1]; ++k) dot += v[k] * x[ci[k]];
    y[r] = dot;
  }
}

int vclo_mixed_cg(int rows, const u32 *rp, const u32 *ci, const double *v, const double *b, double *x,
                  double tol, int maxit, float inner_tol, int *iters, double *err, int *outer_updates)
{
  size_t n = (size_t)rows, nnz = rp[rows];
  double *res = (double*)malloc(sizeof(double) * n);
  float *vl = (float*)malloc(sizeof(float) * (nnz ? nnz : 1));
  float *rl = (float*)malloc(sizeof(float) * n), *xl = (float*)calloc(n, sizeof(float));
  float *pl = (float*)malloc(sizeof(float) * n), *tl = (float*)malloc(sizeof(float) * n);
  memset(x, 0, sizeof(double) * n);
  *iters = 0; *err = 0; if (outer_updates) *outer_updates = 0;
  double ip_rr = 0;
  for (size_t i = 0; i < n; ++i) ip_rr += b[i] * b[i];
  double new_ip_rr = 0, norm_rhs_squared = ip_rr;
  if (norm_rhs_squared <= 0) goto done;
  for (size_t k = 0; k < nnz; ++k) vl[k] = (float)v[k];
  for (size_t i = 0; i < n; ++i) { pl[i] = (float)b[i]; rl[i] = pl[i]; }
  float inner_ip_rr = (float)ip_rr, initial_inner = (float)ip_rr;

  for (int i = 0; i < maxit; ++i)
  {
    *iters = i + 1;
    spmv_f(rows, rp, ci, vl, pl, tl);
    float tp = 0;
    for (size_t k = 0; k < n; ++k) tp += tl[k] * pl[k];
    float alpha = inner_ip_rr / tp;
    float new_inner = 0;
    for (size_t k = 0; k < n; ++k) { xl[k] += alpha * pl[k]; rl[k] -= alpha * tl[k]; }
    for (size_t k = 0; k < n; ++k) new_inner += rl[k] * rl[k];
    float beta = new_inner / inner_ip_rr;
    inner_ip_rr = new_inner;
    for (size_t k = 0; k < n; ++k) pl[k] = rl[k] + beta * pl[k];

    if (new_inner < inner_tol * initial_inner || i == maxit - 1)
    {
      if (outer_updates) (*outer_updates)++;
      for (size_t k = 0; k < n; ++k) x[k] += (double)xl[k];
      spmv_d(rows, rp, ci, v, x, res);
      new_ip_rr = 0;
      for (size_t k = 0; k < n; ++k) { res[k] = b[k] - res[k]; new_ip_rr += res[k] * res[k]; }
      if (new_ip_rr / norm_rhs_squared < tol * tol) break;
      for (size_t k = 0; k < n; ++k) { pl[k] = (float)res[k]; rl[k] = pl[k]; xl[k] = 0; }
      initial_inner = (float)new_ip_rr;
      inner_ip_rr = (float)new_ip_rr;
    }
  }
  *err = sqrt(new_ip_rr / norm_rhs_squared);
done:
  free(res); free(vl); free(rl); free(xl); free(pl); free(tl);
  return 0;
}
