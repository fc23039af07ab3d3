/* examples/c_api_cg.c -- the C-ABI of libvcl_b200.so used from plain C (C99, no CUDA toolkit, no C++): what a C / Fortran /
 * Julia binding of the library looks like, in the style of libviennacl's C interface (libviennacl/include/viennacl.hpp).
 *
 *   gcc -std=c99 -Iinclude examples/c_api_cg.c -Lviennacl-dev_b200/lib -lvcl_b200 -Wl,-rpath,$PWD/viennacl-dev_b200/lib -lm -o c_api_cg
 *
 * Builds the 2-D 5-point Laplacian on the device, computes y = A x and solves A u = 1 with CG, Jacobi-PCG and (single
 * precision) CG, reading the results back through the ABI's own memory functions. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "vcl_b200.h"
#include "vcl_b200_float.h"

#define CHECK(call) do { ViennaCLStatus st_ = (call); if (st_ != ViennaCLSuccess) { \
  fprintf(stderr, "%s failed with status %d: %s\n", #call, (int)st_, backend ? ViennaCLBackendLastError(backend) : "no backend"); return EXIT_FAILURE; } } while (0)

int main(void)
{
  ViennaCLBackend backend = NULL;
  const int nx = 200, ny = 150;
  long long rows = 0, nnz = 0;
  unsigned int *row_ptr = NULL, *col_idx = NULL, *row_blocks = NULL;
  double *values = NULL, *b = NULL, *u = NULL, *y = NULL;
  double *host = NULL, nrm = 0.0;
  ViennaCLInt num_blocks = 0;
  int i;

  if (ViennaCLBackendCreate(&backend) != ViennaCLSuccess)
  {
    fprintf(stderr, "no usable sm_100 device: libvcl_b200 has no CPU fallback\n");
    return EXIT_FAILURE;
  }
  printf("%s\n", ViennaCLB200Version());

  /* the matrix: size query, allocation, generation (tools/matrix_generation.hpp:47-88 semantics) */
  CHECK(ViennaCLCUDADgenerate_stencil(backend, nx, ny, 1, 0.0, 0.0, 0.0, NULL, NULL, NULL, &rows, &nnz));
  CHECK(ViennaCLCUDAMemAlloc(backend, (void**)&row_ptr, sizeof(unsigned int) * (size_t)(rows + 1)));
  CHECK(ViennaCLCUDAMemAlloc(backend, (void**)&col_idx, sizeof(unsigned int) * (size_t)nnz));
  CHECK(ViennaCLCUDAMemAlloc(backend, (void**)&values, sizeof(double) * (size_t)nnz));
  CHECK(ViennaCLCUDADgenerate_stencil(backend, nx, ny, 1, 0.0, 0.0, 0.0, row_ptr, col_idx, values, &rows, &nnz));
  /* row blocks (compressed_matrix::generate_row_block_information): count, then fill */
  CHECK(ViennaCLCUDAcsr_row_blocks(backend, (ViennaCLInt)rows, row_ptr, NULL, &num_blocks));
  CHECK(ViennaCLCUDAMemAlloc(backend, (void**)&row_blocks, sizeof(unsigned int) * (size_t)(num_blocks + 1)));
  CHECK(ViennaCLCUDAcsr_row_blocks(backend, (ViennaCLInt)rows, row_ptr, row_blocks, &num_blocks));

  CHECK(ViennaCLCUDAMemAlloc(backend, (void**)&b, sizeof(double) * (size_t)rows));
  CHECK(ViennaCLCUDAMemAlloc(backend, (void**)&u, sizeof(double) * (size_t)rows));
  CHECK(ViennaCLCUDAMemAlloc(backend, (void**)&y, sizeof(double) * (size_t)rows));
  CHECK(ViennaCLCUDADassign(backend, (ViennaCLInt)rows, b, 0, 1, 1.0));

  /* y = A * 1: interior rows sum to 0, boundary rows to the number of missing neighbours */
  CHECK(ViennaCLCUDADcsrmv(backend, (ViennaCLInt)rows, (ViennaCLInt)rows, (ViennaCLInt)nnz, row_ptr, col_idx, values, row_blocks, num_blocks,
                           b, 0, 1, 1.0, y, 0, 1, 0.0));
  host = (double*)malloc(sizeof(double) * (size_t)rows);
  CHECK(ViennaCLCUDAMemRead(backend, y, 0, host, sizeof(double) * (size_t)rows, 0));
  printf("A*1: corner %.1f, edge %.1f, interior %.1f\n", host[0], host[1], host[nx + 1]);
  if (host[0] != 2.0 || host[1] != 1.0 || host[nx + 1] != 0.0) { fprintf(stderr, "unexpected product\n"); return EXIT_FAILURE; }

  /* solve A u = 1: pipelined CG, then CG with the Jacobi preconditioner fused on the device */
  for (i = 0; i < 2; ++i)
  {
    ViennaCLCUDADcsr A;
    ViennaCLB200SolverTag tag;
    A.rows = (ViennaCLInt)rows; A.cols = (ViennaCLInt)rows; A.nnz = (ViennaCLInt)nnz;
    A.row_ptr = row_ptr; A.col_idx = col_idx; A.values = values; A.row_blocks = row_blocks; A.num_blocks = num_blocks;
    tag.tolerance = 1e-9; tag.abs_tolerance = 0.0; tag.max_iterations = 2000; tag.krylov_dim = 0; tag.max_iterations_before_restart = 0;
    tag.precond = i == 0 ? ViennaCLB200PrecondNone : ViennaCLB200PrecondJacobi;
    tag.monitor = NULL; tag.monitor_user = NULL; tag.iters = 0; tag.error = 0.0;
    CHECK(ViennaCLCUDADcsr_cg(backend, &A, b, u, &tag));
    /* true residual: y = b - A u */
    CHECK(ViennaCLCUDADav(backend, (ViennaCLInt)rows, y, 0, 1, b, 0, 1, 1.0));
    CHECK(ViennaCLCUDADcsrmv(backend, (ViennaCLInt)rows, (ViennaCLInt)rows, (ViennaCLInt)nnz, row_ptr, col_idx, values, row_blocks, num_blocks,
                             u, 0, 1, -1.0, y, 0, 1, 1.0));
    CHECK(ViennaCLCUDADnrm2(backend, (ViennaCLInt)rows, &nrm, y, 0, 1));
    printf("%s: %d iterations, estimate %.2e, true relative residual %.2e\n", i == 0 ? "CG" : "CG + Jacobi", (int)tag.iters, tag.error,
           nrm / sqrt((double)rows));
    if (!(tag.error < 1e-9) || !(nrm / sqrt((double)rows) < 1e-8)) { fprintf(stderr, "solver did not converge\n"); return EXIT_FAILURE; }
  }

  /* the same solve in single precision through the S entry points (vcl_b200_float.h) */
  {
    float *values_f = NULL, *b_f = NULL, *u_f = NULL;
    ViennaCLCUDAScsr A;
    ViennaCLB200SolverTagS tag;
    CHECK(ViennaCLCUDAMemAlloc(backend, (void**)&values_f, sizeof(float) * (size_t)nnz));
    CHECK(ViennaCLCUDAMemAlloc(backend, (void**)&b_f, sizeof(float) * (size_t)rows));
    CHECK(ViennaCLCUDAMemAlloc(backend, (void**)&u_f, sizeof(float) * (size_t)rows));
    CHECK(ViennaCLCUDAconvert_DtoS(backend, nnz, values, values_f));
    CHECK(ViennaCLCUDAconvert_DtoS(backend, rows, b, b_f));
    A.rows = (ViennaCLInt)rows; A.cols = (ViennaCLInt)rows; A.nnz = (ViennaCLInt)nnz;
    A.row_ptr = row_ptr; A.col_idx = col_idx; A.values = values_f; A.row_blocks = row_blocks; A.num_blocks = num_blocks;
    tag.tolerance = 1e-5; tag.abs_tolerance = 0.0; tag.max_iterations = 2000; tag.krylov_dim = 0; tag.max_iterations_before_restart = 0;
    tag.precond = ViennaCLB200PrecondNone; tag.monitor = NULL; tag.monitor_user = NULL; tag.iters = 0; tag.error = 0.0;
    CHECK(ViennaCLCUDAScsr_cg(backend, &A, b_f, u_f, &tag));
    printf("CG (float): %d iterations, estimate %.2e\n", (int)tag.iters, tag.error);
    if (!(tag.error < 1e-5)) { fprintf(stderr, "float solver did not converge\n"); return EXIT_FAILURE; }
    CHECK(ViennaCLCUDAMemFree(backend, values_f)); CHECK(ViennaCLCUDAMemFree(backend, b_f)); CHECK(ViennaCLCUDAMemFree(backend, u_f));
  }

  free(host);
  CHECK(ViennaCLCUDAMemFree(backend, row_ptr)); CHECK(ViennaCLCUDAMemFree(backend, col_idx)); CHECK(ViennaCLCUDAMemFree(backend, values));
  CHECK(ViennaCLCUDAMemFree(backend, row_blocks)); CHECK(ViennaCLCUDAMemFree(backend, b)); CHECK(ViennaCLCUDAMemFree(backend, u));
  CHECK(ViennaCLCUDAMemFree(backend, y));
  CHECK(ViennaCLBackendDestroy(&backend));
  printf("!!!! C EXAMPLE COMPLETED SUCCESSFULLY !!!!\n");
  return EXIT_SUCCESS;
}
