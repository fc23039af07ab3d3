// viennacl/linalg/mixed_precision_cg.hpp -- solve(A, b, mixed_precision_cg_tag) for compressed_matrix<double>
// (reference: linalg/mixed_precision_cg.hpp:46-201).  The whole solve is one C-ABI call
// (ViennaCLCUDADcsr_mixed_precision_cg): float inner iterations on the fused single-precision pipelined CG kernels
// (8 instead of 12 bytes per matrix entry), double residual restarts.
#ifndef VIENNACL_B200_LINALG_MIXED_PRECISION_CG_HPP
#define VIENNACL_B200_LINALG_MIXED_PRECISION_CG_HPP
#include <cassert>
#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/linalg/cg.hpp"
namespace viennacl
{
namespace linalg
{
/** @brief Tag of the mixed-precision CG solver; same parameters, defaults and accessors as the reference (:46-87) */
class mixed_precision_cg_tag
{
public:
  mixed_precision_cg_tag(double tol = 1e-8, unsigned int max_iterations = 300, float inner_tol = 1e-2f)
    : tol_(tol), iterations_(max_iterations), inner_tol_(inner_tol), iters_taken_(0), last_error_(0) {}
  double tolerance() const { return tol_; }
  float inner_tolerance() const { return inner_tol_; }
  unsigned int max_iterations() const { return iterations_; }
  unsigned int iters() const { return iters_taken_; }
  void iters(unsigned int i) const { iters_taken_ = i; }
  double error() const { return last_error_; }
  void error(double e) const { last_error_ = e; }
private:
  double tol_;
  unsigned int iterations_;
  float inner_tol_;
  mutable unsigned int iters_taken_;
  mutable double last_error_;
};

/** @brief x = solve(A, b, mixed_precision_cg_tag(...)), :95-186.  The float copy of the matrix values lives in backend
 *  workspace for the duration of the call; the index arrays and row blocks of A are shared, not copied (the reference
 *  copies them, :131-132). */
template<unsigned int AlignmentV>
viennacl::vector<double> solve(compressed_matrix<double, AlignmentV> const & A, vector_base<double> const & rhs, mixed_precision_cg_tag const & tag)
{
  assert(A.size1() == rhs.size() && A.size1() == A.size2() && bool("solve() needs a square system of matching size"));
  viennacl::vector<double> result(rhs.size());
  viennacl::vector<double> compact;
  const vector_base<double> *b = &rhs;
  if (rhs.stride() != 1) { compact = rhs; b = &compact; }
  ViennaCLB200SolverTag t = detail::to_abi(cg_tag(tag.tolerance(), tag.max_iterations()));
  ViennaCLCUDADcsr a = A.abi();
  backend::b200::check(ViennaCLCUDADcsr_mixed_precision_cg(backend::b200::handle(), &a, NULL, b->ptr() + b->start(), result.ptr(),
                                                           tag.inner_tolerance(), &t));
  tag.iters(static_cast<unsigned int>(t.iters)); tag.error(t.error);
  return result;
}

template<unsigned int AlignmentV>
viennacl::vector<double> solve(compressed_matrix<double, AlignmentV> const & A, vector_base<double> const & rhs, mixed_precision_cg_tag const & tag,
                               viennacl::linalg::no_precond)
{ return solve(A, rhs, tag); }
}
}
#endif
