// viennacl/linalg/bicgstab.hpp -- pipelined BiCGStab solver (+ fused Jacobi-preconditioned variant) (reference: linalg/bicgstab.hpp:47-90, 97-215, 398-489, 495-592).
// The tag keeps the reference's fields and defaults; solve() forwards to the whole-solve entry point of the C-ABI
// (ViennaCLCUDAD{csr,sell}_bicgstab), whose loop runs next to the kernels with device-resident scalars.
#ifndef VIENNACL_B200_LINALG_BICGSTAB_HPP
#define VIENNACL_B200_LINALG_BICGSTAB_HPP
#include "viennacl/linalg/detail_solver_call.hpp"
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

  template<typename MatrixT, typename NumericT>
  viennacl::vector<NumericT> solve_impl(MatrixT const & A, vector_base<NumericT> const & rhs, bicgstab_tag const & tag, viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  {
    ViennaCLB200SolverTag t = to_abi(tag);
    viennacl::vector<NumericT> x = run(SOLVER_BICGSTAB, A, rhs, t, monitor, monitor_data);
    tag.iters(static_cast<unsigned int>(t.iters)); tag.error(t.error);
    return x;
  }

  /** @brief Left-preconditioned BiCGStab with Jacobi (bicgstab.hpp:398-489): fused on the device, 5 kernels per iteration */
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(compressed_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, bicgstab_tag const & tag,
                                        jacobi_precond< compressed_matrix<NumericT, AlignmentV> > const &,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  {
    ViennaCLB200SolverTag t = to_abi(tag);
    t.precond = ViennaCLB200PrecondJacobi;
    viennacl::vector<NumericT> x = run(SOLVER_BICGSTAB, A, rhs, t, monitor, monitor_data);
    tag.iters(static_cast<unsigned int>(t.iters)); tag.error(t.error);
    return x;
  }
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
#endif
