// viennacl/linalg/cg.hpp -- pipelined conjugate gradient solver (reference: linalg/cg.hpp:48-87, 128-187, 338-434).
// The tag keeps the reference's fields and defaults; solve() forwards to the whole-solve entry point of the C-ABI
// (ViennaCLCUDAD{csr,sell}_cg), whose loop runs next to the kernels with device-resident scalars.
#ifndef VIENNACL_B200_LINALG_CG_HPP
#define VIENNACL_B200_LINALG_CG_HPP
#include <cmath>
#include "viennacl/linalg/detail_solver_call.hpp"
#include "viennacl/linalg/inner_prod.hpp"
#include "viennacl/linalg/norm_2.hpp"
namespace viennacl
{
namespace linalg
{

/** @brief Solver configuration and result carrier; iters()/error() are mutable so that a const tag reports back (cg.hpp:48-87) */
class cg_tag
{
public:
  cg_tag(double tol = 1e-8, unsigned int max_iterations = 300) : tol_(tol), abs_tol_(0), iterations_(max_iterations), iters_taken_(0), last_error_(0) {}
  double tolerance() const { return tol_; }
  double abs_tolerance() const { return abs_tol_; }
  void abs_tolerance(double new_tol) { if (new_tol >= 0) abs_tol_ = new_tol; }
  unsigned int max_iterations() const { return iterations_; }

  unsigned int iters() const { return iters_taken_; }
  void iters(unsigned int i) const { iters_taken_ = i; }
  double error() const { return last_error_; }
  void error(double e) const { last_error_ = e; }
private:
  double tol_;
  double abs_tol_;
  unsigned int iterations_;

  mutable unsigned int iters_taken_;
  mutable double last_error_;
};

namespace detail
{
  inline ViennaCLB200SolverTag to_abi(cg_tag const & tag)
  {
    ViennaCLB200SolverTag t;
    t.tolerance = tag.tolerance(); t.abs_tolerance = tag.abs_tolerance(); t.max_iterations = ViennaCLInt(tag.max_iterations());
    t.krylov_dim = 0; t.max_iterations_before_restart = 0; t.precond = ViennaCLB200PrecondNone;
    t.monitor = NULL; t.monitor_user = NULL; t.iters = 0; t.error = 0;
    return t;
  }

  /** @brief Pipelined CG on the device (cg.hpp:128-187): compressed_matrix / sliced_ell_matrix without preconditioner */
  template<typename MatrixT, typename NumericT>
  viennacl::vector<NumericT> fused_cg(MatrixT const & A, vector_base<NumericT> const & rhs, cg_tag const & tag,
                                      bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*), void *monitor_data,
                                      ViennaCLB200Precond pc = ViennaCLB200PrecondNone)
  {
    ViennaCLB200SolverTag t = to_abi(tag);
    t.precond = pc;
    viennacl::vector<NumericT> x = run(SOLVER_CG, A, rhs, t, monitor, monitor_data);
    tag.iters(static_cast<unsigned int>(t.iters)); tag.error(t.error);
    return x;
  }
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(compressed_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, cg_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_cg(A, rhs, tag, monitor, monitor_data); }
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(ell_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, cg_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_cg(A, rhs, tag, monitor, monitor_data); }
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(hyb_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, cg_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_cg(A, rhs, tag, monitor, monitor_data); }
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(coordinate_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, cg_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_cg(A, rhs, tag, monitor, monitor_data); }
  template<typename NumericT, typename IndexT>
  viennacl::vector<NumericT> solve_impl(sliced_ell_matrix<NumericT, IndexT> const & A, vector_base<NumericT> const & rhs, cg_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_cg(A, rhs, tag, monitor, monitor_data); }

  /** @brief CG with the Jacobi preconditioner on a compressed_matrix: fused single-reduction PCG on the device (2 kernels per
   *  iteration) instead of the reference's generic path (cg.hpp:257-322) */
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(compressed_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, cg_tag const & tag,
                                        jacobi_precond< compressed_matrix<NumericT, AlignmentV> > const &,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_cg(A, rhs, tag, monitor, monitor_data, ViennaCLB200PrecondJacobi); }

  /** @brief CG with a row_scaling preconditioner (row_scaling.hpp:150-190) on a compressed_matrix: same fused path, the scaling
   *  vector holds row norms instead of the diagonal */
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(compressed_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, cg_tag const & tag,
                                        row_scaling< compressed_matrix<NumericT, AlignmentV> > const & precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_cg(A, rhs, tag, monitor, monitor_data, precond.abi_id()); }

  /** @brief Preconditioned CG for ANY operator (matrix-free `apply()`) and ANY preconditioner with `apply(v)`:
   *  the reference's generic path (cg.hpp:257-322; Saad, Alg. 9.1), built from prod / inner_prod / vector expressions.
   *  One blocking reduction per inner product, exactly like the reference -- the fused paths above avoid that. */
  template<typename MatrixT, typename NumericT, typename PreconditionerT>
  viennacl::vector<NumericT> solve_impl(MatrixT const & A, vector_base<NumericT> const & rhs, cg_tag const & tag, PreconditionerT const & precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  {
    typedef viennacl::vector<NumericT> VectorT;
    VectorT result(rhs.size());
    VectorT residual = rhs;
    VectorT z = rhs;
    precond.apply(z);
    VectorT p = z;
    VectorT Ap(rhs.size());
    NumericT ip_rz = viennacl::linalg::inner_prod(residual, z);
    const NumericT norm_rhs_squared = ip_rz;
    NumericT new_ip_rz = 0;
    tag.iters(0); tag.error(0);
    if (std::fabs(norm_rhs_squared) <= tag.abs_tolerance() * tag.abs_tolerance()) return result;

    for (unsigned int i = 0; i < tag.max_iterations(); ++i)
    {
      tag.iters(i + 1);
      Ap = viennacl::linalg::prod(A, p);
      const NumericT alpha = ip_rz / NumericT(viennacl::linalg::inner_prod(Ap, p));
      result += alpha * p;
      residual -= alpha * Ap;
      z = residual;
      precond.apply(z);
      new_ip_rz = viennacl::linalg::inner_prod(residual, z);
      const NumericT rel_sq = new_ip_rz / norm_rhs_squared;
      if (monitor && monitor(result, std::sqrt(std::fabs(rel_sq)), monitor_data)) break;
      if (std::fabs(rel_sq) < tag.tolerance() * tag.tolerance() || std::fabs(new_ip_rz) < tag.abs_tolerance() * tag.abs_tolerance()) break;
      const NumericT beta = new_ip_rz / ip_rz;
      ip_rz = new_ip_rz;
      p = z + beta * p;
    }
    tag.error(std::sqrt(std::fabs(new_ip_rz / norm_rhs_squared)));
    return result;
  }

}

/** @brief x = solve(A, b, cg_tag(...)) for compressed_matrix / sliced_ell_matrix (cg.hpp:338-375) */
template<typename MatrixT, typename NumericT, typename PreconditionerT>
viennacl::vector<NumericT> solve(MatrixT const & A, vector_base<NumericT> const & rhs, cg_tag const & tag, PreconditionerT const & precond)
{ return detail::solve_impl(A, rhs, tag, precond); }

template<typename MatrixT, typename NumericT>
viennacl::vector<NumericT> solve(MatrixT const & A, vector_base<NumericT> const & rhs, cg_tag const & tag)
{ return detail::solve_impl(A, rhs, tag, viennacl::linalg::no_precond()); }

/** @brief Functor form with initial guess and monitor (cg.hpp:379-434) */
template<typename VectorT>
class cg_solver
{
public:
  typedef typename VectorT::value_type numeric_type;

  cg_solver(cg_tag const & tag) : tag_(tag), monitor_callback_(NULL), user_data_(NULL) {}

  template<typename MatrixT, typename PreconditionerT>
  VectorT operator()(MatrixT const & A, VectorT const & b, PreconditionerT const & precond) const
  {
    if (viennacl::traits::size(init_guess_) > 0)          // A y = b - A x0, x = x0 + y
    {
      VectorT mod_rhs = viennacl::linalg::prod(A, init_guess_);
      mod_rhs = b - mod_rhs;
      VectorT y = detail::solve_impl(A, mod_rhs, tag_, precond, monitor_callback_, user_data_);
      VectorT x = init_guess_ + y;
      return x;
    }
    return detail::solve_impl(A, b, tag_, precond, monitor_callback_, user_data_);
  }

  template<typename MatrixT>
  VectorT operator()(MatrixT const & A, VectorT const & b) const { return operator()(A, b, viennacl::linalg::no_precond()); }

  void set_initial_guess(VectorT const & x) { init_guess_ = x; }
  void set_monitor(bool (*monitor_fun)(VectorT const &, numeric_type, void *), void *user_data) { monitor_callback_ = monitor_fun; user_data_ = user_data; }
  cg_tag const & tag() const { return tag_; }

private:
  cg_tag tag_;
  VectorT init_guess_;
  bool (*monitor_callback_)(VectorT const &, numeric_type, void *);
  void *user_data_;
};

}
}
#include "viennacl/linalg/stl_solve.hpp"
#endif
