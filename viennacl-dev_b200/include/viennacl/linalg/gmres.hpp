// viennacl/linalg/gmres.hpp -- restarted GMRES(m), pipelined simpler-GMRES with fused classical Gram-Schmidt (reference: linalg/gmres.hpp:49-101, 181-367, 635-732).
// The tag keeps the reference's fields and defaults; solve() forwards to the whole-solve entry point of the C-ABI
// (ViennaCLCUDAD{csr,sell}_gmres), whose loop runs next to the kernels with device-resident scalars.
#ifndef VIENNACL_B200_LINALG_GMRES_HPP
#define VIENNACL_B200_LINALG_GMRES_HPP
#include <cmath>
#include <vector>
#include <algorithm>
#include "viennacl/linalg/detail_solver_call.hpp"
#include "viennacl/linalg/inner_prod.hpp"
#include "viennacl/linalg/norm_2.hpp"
namespace viennacl
{
namespace linalg
{

/** @brief Solver configuration and result carrier; iters()/error() are mutable so that a const tag reports back (gmres.hpp:49-101) */
class gmres_tag
{
public:
  gmres_tag(double tol = 1e-10, unsigned int max_iterations = 300, unsigned int krylov_dim = 20)
    : tol_(tol), abs_tol_(0), iterations_(max_iterations), krylov_dim_(krylov_dim), iters_taken_(0), last_error_(0) {}
  double tolerance() const { return tol_; }
  double abs_tolerance() const { return abs_tol_; }
  void abs_tolerance(double new_tol) { if (new_tol >= 0) abs_tol_ = new_tol; }
  unsigned int max_iterations() const { return iterations_; }
  unsigned int krylov_dim() const { return krylov_dim_; }
  unsigned int max_restarts() const
  {
    unsigned int ret = iterations_ / krylov_dim_;
    if (ret > 0 && (ret * krylov_dim_ == iterations_)) return ret - 1;
    return ret;
  }
  unsigned int iters() const { return iters_taken_; }
  void iters(unsigned int i) const { iters_taken_ = i; }
  double error() const { return last_error_; }
  void error(double e) const { last_error_ = e; }
private:
  double tol_;
  double abs_tol_;
  unsigned int iterations_;
  unsigned int krylov_dim_;
  mutable unsigned int iters_taken_;
  mutable double last_error_;
};

namespace detail
{
  inline ViennaCLB200SolverTag to_abi(gmres_tag const & tag)
  {
    ViennaCLB200SolverTag t;
    t.tolerance = tag.tolerance(); t.abs_tolerance = tag.abs_tolerance(); t.max_iterations = ViennaCLInt(tag.max_iterations());
    t.krylov_dim = ViennaCLInt(tag.krylov_dim()); t.max_iterations_before_restart = 0; t.precond = ViennaCLB200PrecondNone;
    t.monitor = NULL; t.monitor_user = NULL; t.iters = 0; t.error = 0;
    return t;
  }

  /** @brief Pipelined GMRES(m) on the device (gmres.hpp:181-367): compressed_matrix / sliced_ell_matrix without preconditioner */
  template<typename MatrixT, typename NumericT>
  viennacl::vector<NumericT> fused_gmres(MatrixT const & A, vector_base<NumericT> const & rhs, gmres_tag const & tag,
                                         bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*), void *monitor_data,
                                         ViennaCLB200Precond pc = ViennaCLB200PrecondNone)
  {
    ViennaCLB200SolverTag t = to_abi(tag);
    t.precond = pc;
    viennacl::vector<NumericT> x = run(SOLVER_GMRES, A, rhs, t, monitor, monitor_data);
    tag.iters(static_cast<unsigned int>(t.iters)); tag.error(t.error);
    return x;
  }
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(compressed_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, gmres_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_gmres(A, rhs, tag, monitor, monitor_data); }
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(ell_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, gmres_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_gmres(A, rhs, tag, monitor, monitor_data); }
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(hyb_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, gmres_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_gmres(A, rhs, tag, monitor, monitor_data); }
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(coordinate_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, gmres_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_gmres(A, rhs, tag, monitor, monitor_data); }
  template<typename NumericT, typename IndexT>
  viennacl::vector<NumericT> solve_impl(sliced_ell_matrix<NumericT, IndexT> const & A, vector_base<NumericT> const & rhs, gmres_tag const & tag,
                                        viennacl::linalg::no_precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_gmres(A, rhs, tag, monitor, monitor_data); }

  /** @brief GMRES(m) with the Jacobi preconditioner on a compressed_matrix: the pipelined cycle on D^-1 A (divide folded into
   *  the SpMV epilogue) with the reference's per-iteration stopping rule (gmres.hpp:579-584) applied to the cycle's
   *  projections -- same iterate and iteration count as the reference's Householder path (:449-631), no blocking reduction
   *  per inner product. */
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(compressed_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, gmres_tag const & tag,
                                        jacobi_precond< compressed_matrix<NumericT, AlignmentV> > const &,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_gmres(A, rhs, tag, monitor, monitor_data, ViennaCLB200PrecondJacobi); }

  /** @brief GMRES(m) with a row_scaling preconditioner (row_scaling.hpp:150-190): same fused path */
  template<typename NumericT, unsigned int AlignmentV>
  viennacl::vector<NumericT> solve_impl(compressed_matrix<NumericT, AlignmentV> const & A, vector_base<NumericT> const & rhs, gmres_tag const & tag,
                                        row_scaling< compressed_matrix<NumericT, AlignmentV> > const & precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  { return fused_gmres(A, rhs, tag, monitor, monitor_data, precond.abi_id()); }

  /** @brief Left-preconditioned restarted GMRES(m) for ANY operator and ANY preconditioner with `apply(v)`.
   *  Same problem statement, stopping rule and bookkeeping as the reference's generic path (gmres.hpp:449-631): the residual
   *  M^-1 (b - A x) is minimised over the Krylov space of M^-1 A, the estimate |rho * rho_0| / ||b|| is tested after every
   *  inner iteration, tag.iters() counts inner iterations, the monitor runs once per restart.  The reference orthogonalises
   *  with Householder reflections; here the basis is built by modified Gram-Schmidt and the least-squares problem is kept
   *  triangular with Givens rotations -- the same minimiser, so iteration counts agree to within rounding. */
  template<typename MatrixT, typename NumericT, typename PreconditionerT>
  viennacl::vector<NumericT> solve_impl(MatrixT const & A, vector_base<NumericT> const & rhs, gmres_tag const & tag, PreconditionerT const & precond,
                                        bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*) = NULL, void *monitor_data = NULL)
  {
    typedef viennacl::vector<NumericT> VectorT;
    const vcl_size_t n = rhs.size();
    vcl_size_t m = tag.krylov_dim();
    if (n < m) m = n;
    VectorT result(n), w(n);
    const NumericT norm_rhs = viennacl::linalg::norm_2(rhs);
    tag.iters(0); tag.error(0);
    if (norm_rhs <= tag.abs_tolerance() || m == 0) return result;

    std::vector<VectorT> V(m + 1);
    std::vector< std::vector<NumericT> > H(m + 1, std::vector<NumericT>(m, NumericT(0)));   // H[i][k], upper triangular after rotations
    std::vector<NumericT> cs(m), sn(m), g(m + 1), y(m);

    for (unsigned int it = 0; it <= tag.max_restarts(); ++it)
    {
      w = rhs - viennacl::linalg::prod(A, result);
      precond.apply(w);
      const NumericT rho_0 = viennacl::linalg::norm_2(w);
      if (rho_0 / norm_rhs < tag.tolerance() || rho_0 < tag.abs_tolerance()) { tag.error(rho_0 / norm_rhs); return result; }
      V[0] = (NumericT(1) / rho_0) * w;
      std::fill(g.begin(), g.end(), NumericT(0));
      g[0] = rho_0;
      NumericT estimate = rho_0 / norm_rhs;

      vcl_size_t k = 0;
      for (k = 0; k < m; ++k)
      {
        tag.iters(tag.iters() + 1);
        w = viennacl::linalg::prod(A, V[k]);
        precond.apply(w);
        for (vcl_size_t i = 0; i <= k; ++i)                  // modified Gram-Schmidt
        {
          H[i][k] = viennacl::linalg::inner_prod(w, V[i]);
          w -= H[i][k] * V[i];
        }
        const NumericT h_next = viennacl::linalg::norm_2(w);
        for (vcl_size_t i = 0; i < k; ++i)                   // earlier rotations on the new column
        {
          const NumericT t = cs[i] * H[i][k] + sn[i] * H[i + 1][k];
          H[i + 1][k] = -sn[i] * H[i][k] + cs[i] * H[i + 1][k];
          H[i][k] = t;
        }
        const NumericT denom = std::sqrt(H[k][k] * H[k][k] + h_next * h_next);
        cs[k] = denom > 0 ? H[k][k] / denom : NumericT(1);
        sn[k] = denom > 0 ? h_next / denom : NumericT(0);
        H[k][k] = denom;
        g[k + 1] = -sn[k] * g[k];
        g[k] = cs[k] * g[k];
        estimate = std::fabs(g[k + 1]) / norm_rhs;
        if (h_next > 0 && k + 1 < m + 1) V[k + 1] = (NumericT(1) / h_next) * w;
        if (estimate < tag.tolerance() || h_next <= 0) { ++k; break; }
      }

      for (vcl_size_t i2 = k; i2 > 0; --i2)                  // back substitution on the k x k triangle
      {
        const vcl_size_t i = i2 - 1;
        NumericT sum = g[i];
        for (vcl_size_t j = i + 1; j < k; ++j) sum -= H[i][j] * y[j];
        y[i] = sum / H[i][i];
      }
      for (vcl_size_t i = 0; i < k; ++i) result += y[i] * V[i];

      tag.error(estimate);
      if (monitor && monitor(result, estimate, monitor_data)) break;
      if (tag.error() < tag.tolerance()) return result;
    }
    return result;
  }

}

/** @brief x = solve(A, b, gmres_tag(...)) for compressed_matrix / sliced_ell_matrix (gmres.hpp:635-673) */
template<typename MatrixT, typename NumericT, typename PreconditionerT>
viennacl::vector<NumericT> solve(MatrixT const & A, vector_base<NumericT> const & rhs, gmres_tag const & tag, PreconditionerT const & precond)
{ return detail::solve_impl(A, rhs, tag, precond); }

template<typename MatrixT, typename NumericT>
viennacl::vector<NumericT> solve(MatrixT const & A, vector_base<NumericT> const & rhs, gmres_tag const & tag)
{ return detail::solve_impl(A, rhs, tag, viennacl::linalg::no_precond()); }

/** @brief Functor form with initial guess and monitor (gmres.hpp:677-732) */
template<typename VectorT>
class gmres_solver
{
public:
  typedef typename VectorT::value_type numeric_type;

  gmres_solver(gmres_tag const & tag) : tag_(tag), monitor_callback_(NULL), user_data_(NULL) {}

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
  gmres_tag const & tag() const { return tag_; }

private:
  gmres_tag tag_;
  VectorT init_guess_;
  bool (*monitor_callback_)(VectorT const &, numeric_type, void *);
  void *user_data_;
};

}
}
#include "viennacl/linalg/stl_solve.hpp"
#endif
