// viennacl/linalg/bicgstab.hpp -- pipelined BiCGStab solver (+ fused Jacobi-preconditioned variant) (reference: linalg/bicgstab.hpp:47-90, 97-215, 398-489, 495-592).
// The tag keeps the reference's fields and defaults; solve() forwards to the whole-solve entry point of the C-ABI
// (ViennaCLCUDAD{csr,sell}_bicgstab), whose loop runs next to the kernels with device-resident scalars.
#ifndef VIENNACL_B200_LINALG_BICGSTAB_HPP
#define VIENNACL_B200_LINALG_BICGSTAB_HPP
#include <cmath>
#include "viennacl/linalg/detail_solver_call.hpp"
#include "viennacl/linalg/inner_prod.hpp"
#include "viennacl/linalg/norm_2.hpp"
namespace viennacl
{
namespace linalg
{

/** @brief Solver configuration and result carrier; iters()/error() are mutable so that a const tag reports back (bicgstab.hpp:47-90) */
class bicgstab_tag
{
public:
  bicgstab_tag(double tol = 1e-8, vcl_size_t max_iters = 400, vcl_size_t max_iters_before_restart = 200)
    : tol_(tol), abs_tol_(0), iterations_(static_cast<unsigned int>(max_iters)), iterations_before_restart_(static_cast<unsigned int>(max_iters_before_restart)), iters_taken_(0), last_error_(0) {}
  double tolerance() const { return tol_; }
  double abs_tolerance() const { return abs_tol_; }
  void abs_tolerance(double new_tol) { if (new_tol >= 0) abs_tol_ = new_tol; }
  unsigned int max_iterations() const { return iterations_; }
  unsigned int max_iterations_before_restart() const { return iterations_before_restart_; }
  unsigned int iters() const { return iters_taken_; }
  void iters(unsigned int i) const { iters_taken_ = i; }
  double error() const { return last_error_; }
  void error(double e) const { last_error_ = e; }
private:
  double tol_;
  double abs_tol_;
  unsigned int iterations_;
  unsigned int iterations_before_restart_;
  mutable unsigned int iters_taken_;
  mutable double last_error_;
};

namespace detail
{
  inline ViennaCLB200SolverTag to_abi(bicgstab_tag const & tag)
  {
    ViennaCLB200SolverTag t;
    t.tolerance = tag.tolerance(); t.abs_tolerance = tag.abs_tolerance(); t.max_iterations = ViennaCLInt(tag.max_iterations());
    t.krylov_dim = 0; t.max_iterations_before_restart = ViennaCLInt(tag.max_iterations_before_restart()); t.precond = ViennaCLB200PrecondNone;
    t.monitor = NULL; t.monitor_user = NULL; t.iters = 0; t.error = 0;
    return t;
  }

  /** @brief Pipelined BiCGStab on the device (bicgstab.hpp:97-215): compressed_matrix / sliced_ell_matrix without preconditioner */
  template<typename MatrixT, typename NumericT>
  viennacl::vector<NumericT> fused_bicgstab(MatrixT const & A, vector_base<NumericT> const & rhs, bicgstab_tag const & tag, ViennaCLB200Precond pc,
                                            bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*), void *monitor_data)
  {
    ViennaCLB200SolverTag t = to_abi(tag);
    t.precond = pc;
    viennacl::vector<NumericT> x = run(SOLVER_BICGSTAB, A, rhs, t, monitor, monitor_data);
    tag.iters(static_cast<unsigned int>(t.iters)); tag.error(t.error);
    return x;
  }
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(compressed_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, bicgstab_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_bicgstab(A, rhs, tag, ViennaCLB200PrecondNone, monitor, monitor_data); }
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(ell_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, bicgstab_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_bicgstab(A, rhs, tag, ViennaCLB200PrecondNone, monitor, monitor_data); }
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(hyb_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, bicgstab_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_bicgstab(A, rhs, tag, ViennaCLB200PrecondNone, monitor, monitor_data); }
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(coordinate_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, bicgstab_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_bicgstab(A, rhs, tag, ViennaCLB200PrecondNone, monitor, monitor_data); }
  template<typename NumericT, typename IndexT>
  viennacl::vector<NumericT> solve_impl(sliced_ell_matrix<NumericT, IndexT> const & A, vector_base<NumericT> const & rhs, bicgstab_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_bicgstab(A, rhs, tag, ViennaCLB200PrecondNone, monitor, monitor_data); }

  /** @brief Left-preconditioned BiCGStab for ANY operator and ANY preconditioner with `apply(v)`: the reference's generic
   *  path (bicgstab.hpp:398-489) including its restart rule, built from prod / inner_prod / norm_2 / vector expressions. */
  template<typename MatrixT, typename NumericT, typename PreconditionerT>
  viennacl::vector<NumericT> solve_impl(MatrixT const & A, vector_base<NumericT> const & rhs, bicgstab_tag const & tag, PreconditionerT const & precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  {
    typedef viennacl::vector<NumericT> VectorT;
    const vcl_size_t n = rhs.size();
    VectorT result(n), residual = rhs, r0star = rhs, p = rhs, s(n), t0(n), t1(n);
    const NumericT norm_rhs = viennacl::linalg::norm_2(rhs);
    NumericT ip_rr0star = norm_rhs * norm_rhs, residual_norm = norm_rhs;
    tag.iters(0); tag.error(0);
    if (norm_rhs <= tag.abs_tolerance()) return result;

    bool restart = true;
    unsigned int last_restart = 0;
    for (unsigned int i = 0; i < tag.max_iterations(); ++i)
    {
      if (restart)
      {
        residual = rhs - viennacl::linalg::prod(A, result);
        precond.apply(residual);
        p = residual;
        r0star = residual;
        ip_rr0star = viennacl::linalg::norm_2(residual);
        ip_rr0star *= ip_rr0star;
        restart = false;
        last_restart = i;
      }
      tag.iters(i + 1);
      t0 = viennacl::linalg::prod(A, p);
      precond.apply(t0);
      const NumericT alpha = ip_rr0star / NumericT(viennacl::linalg::inner_prod(t0, r0star));
      s = residual - alpha * t0;
      t1 = viennacl::linalg::prod(A, s);
      precond.apply(t1);
      const NumericT norm_t1 = viennacl::linalg::norm_2(t1);
      const NumericT omega = NumericT(viennacl::linalg::inner_prod(t1, s)) / (norm_t1 * norm_t1);
      result += alpha * p + omega * s;
      residual = s - omega * t1;
      residual_norm = viennacl::linalg::norm_2(residual);
      if (monitor && monitor(result, std::fabs(residual_norm / norm_rhs), monitor_data)) break;
      if (residual_norm / norm_rhs < tag.tolerance() || residual_norm < tag.abs_tolerance()) break;
      const NumericT new_ip = viennacl::linalg::inner_prod(residual, r0star);
      const NumericT beta = new_ip / ip_rr0star * alpha / omega;
      ip_rr0star = new_ip;
      if (ip_rr0star == NumericT(0) || omega == NumericT(0) || i - last_restart > tag.max_iterations_before_restart()) restart = true;
      p -= omega * t0;                    // p = residual + beta * (p - omega * t0)
      p = residual + beta * p;
    }
    tag.error(residual_norm / norm_rhs);
    return result;
  }

  /** @brief Left-preconditioned BiCGStab with Jacobi (bicgstab.hpp:398-489): fused on the device, 5 kernels per iteration */
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(compressed_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, bicgstab_tag const & tag,
                                        jacobi_precond< compressed_matrix<NumericT, AlignmentV> > const &,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_bicgstab(A, rhs, tag, ViennaCLB200PrecondJacobi, monitor, monitor_data); }

  /** @brief Left-preconditioned BiCGStab with a row_scaling preconditioner (row_scaling.hpp:150-190): same fused path */
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(compressed_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, bicgstab_tag const & tag,
                                        row_scaling< compressed_matrix<NumericT, AlignmentV> > const & precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_bicgstab(A, rhs, tag, precond.abi_id(), monitor, monitor_data); }
}

/** @brief x = solve(A, b, bicgstab_tag(...)) for compressed_matrix / sliced_ell_matrix (bicgstab.hpp:495-533) */
template<typename MatrixT, typename NumericT, typename PreconditionerT>
viennacl::vector<NumericT> solve(MatrixT const & A, vector_base<NumericT> const & rhs, bicgstab_tag const & tag, PreconditionerT const & precond)
{ return detail::solve_impl(A, rhs, tag, precond); }

template<typename MatrixT, typename NumericT>
viennacl::vector<NumericT> solve(MatrixT const & A, vector_base<NumericT> const & rhs, bicgstab_tag const & tag)
{ return detail::solve_impl(A, rhs, tag, viennacl::linalg::no_precond()); }

/** @brief Functor form with initial guess and monitor (bicgstab.hpp:537-592) */
template<typename VectorT>
class bicgstab_solver
{
public:
  typedef typename VectorT::value_type numeric_type;

  bicgstab_solver(bicgstab_tag const & tag) : tag_(tag), monitor_callback_(NULL), user_data_(NULL) {}

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
  bicgstab_tag const & tag() const { return tag_; }

private:
  bicgstab_tag tag_;
  VectorT init_guess_;
  bool (*monitor_callback_)(VectorT const &, numeric_type, void *);
  void *user_data_;
};

}
}
#include "viennacl/linalg/stl_solve.hpp"
#endif
