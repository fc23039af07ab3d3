// viennacl/linalg/detail_solver_call.hpp -- shared glue between the solver tags and the whole-solve entry points of the C-ABI.
#ifndef VIENNACL_B200_LINALG_DETAIL_SOLVER_CALL_HPP
#define VIENNACL_B200_LINALG_DETAIL_SOLVER_CALL_HPP
#include "viennacl/backend/abi.hpp"
#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/sliced_ell_matrix.hpp"
#include "viennacl/ell_matrix.hpp"
#include "viennacl/hyb_matrix.hpp"
#include "viennacl/coordinate_matrix.hpp"
#include "viennacl/linalg/prod.hpp"
#include "viennacl/linalg/jacobi_precond.hpp"
#include "viennacl/linalg/row_scaling.hpp"
namespace viennacl
{
namespace linalg
{
namespace detail
{
  enum solver_kind { SOLVER_CG, SOLVER_BICGSTAB, SOLVER_GMRES };

  /** @brief Adapts the reference's monitor signature bool(*)(vector<T> const &, T, void*) (cg.hpp:420-424) to the C callback */
  template<typename NumericT>
  struct monitor_bridge
  {
    bool (*fun)(viennacl::vector<NumericT> const &, NumericT, void*);
    void *user;
    vcl_size_t size;
    static ViennaCLInt trampoline(const NumericT *x_dev, NumericT est, void *self_)
    {
      monitor_bridge *self = static_cast<monitor_bridge*>(self_);
      viennacl::vector<NumericT> view(const_cast<NumericT*>(x_dev), CUDA_MEMORY, self->size);
      return self->fun(view, NumericT(est), self->user) ? 1 : 0;
    }
  };

  /** @brief Whole-solve entry points of one precision (abi<NumericT>), selected by the matrix struct type */
  template<typename NumericT>
  struct solver_calls
  {
    typedef viennacl::backend::b200::abi<NumericT> abi;
    typedef typename abi::solver_tag tag_type;
    static ViennaCLStatus call(solver_kind k, typename abi::csr const & A, const NumericT *b, NumericT *x, tag_type *t)
    {
      ViennaCLBackend h = backend::b200::handle();
      if (k == SOLVER_CG) return abi::csr_cg(h, &A, b, x, t);
      if (k == SOLVER_BICGSTAB) return abi::csr_bicgstab(h, &A, b, x, t);
      return abi::csr_gmres(h, &A, b, x, t);
    }
    static ViennaCLStatus call(solver_kind k, typename abi::sell const & A, const NumericT *b, NumericT *x, tag_type *t)
    {
      ViennaCLBackend h = backend::b200::handle();
      if (k == SOLVER_CG) return abi::sell_cg(h, &A, b, x, t);
      if (k == SOLVER_BICGSTAB) return abi::sell_bicgstab(h, &A, b, x, t);
      return abi::sell_gmres(h, &A, b, x, t);
    }
    static ViennaCLStatus call(solver_kind k, typename abi::ell const & A, const NumericT *b, NumericT *x, tag_type *t)
    {
      ViennaCLBackend h = backend::b200::handle();
      if (k == SOLVER_CG) return abi::ell_cg(h, &A, b, x, t);
      if (k == SOLVER_BICGSTAB) return abi::ell_bicgstab(h, &A, b, x, t);
      return abi::ell_gmres(h, &A, b, x, t);
    }
    static ViennaCLStatus call(solver_kind k, typename abi::hyb const & A, const NumericT *b, NumericT *x, tag_type *t)
    {
      ViennaCLBackend h = backend::b200::handle();
      if (k == SOLVER_CG) return abi::hyb_cg(h, &A, b, x, t);
      if (k == SOLVER_BICGSTAB) return abi::hyb_bicgstab(h, &A, b, x, t);
      return abi::hyb_gmres(h, &A, b, x, t);
    }
  };

  /** @brief Runs one solve on the device; rhs may be a strided view (it is compacted first). */
  template<typename MatrixT, typename NumericT>
  viennacl::vector<NumericT> run(solver_kind kind, MatrixT const & A, vector_base<NumericT> const & rhs, ViennaCLB200SolverTag & t,
                                 bool (*monitor)(viennacl::vector<NumericT> const &, NumericT, void*), void *monitor_data)
  {
    assert(A.size1() == rhs.size() && A.size1() == A.size2() && bool("solve() needs a square system of matching size"));
    viennacl::vector<NumericT> result(rhs.size());
    viennacl::vector<NumericT> compact;
    const vector_base<NumericT> *b = &rhs;
    if (rhs.stride() != 1) { compact = rhs; b = &compact; }
    monitor_bridge<NumericT> bridge = {monitor, monitor_data, rhs.size()};
    // the tag of this precision: same fields, tolerances and results in double, monitor estimate in NumericT
    typename solver_calls<NumericT>::tag_type tn;
    tn.tolerance = t.tolerance; tn.abs_tolerance = t.abs_tolerance; tn.max_iterations = t.max_iterations; tn.krylov_dim = t.krylov_dim;
    tn.max_iterations_before_restart = t.max_iterations_before_restart; tn.precond = t.precond; tn.iters = 0; tn.error = 0;
    if (monitor) { tn.monitor = &monitor_bridge<NumericT>::trampoline; tn.monitor_user = &bridge; }
    else { tn.monitor = NULL; tn.monitor_user = NULL; }
    backend::b200::check(solver_calls<NumericT>::call(kind, A.abi(), b->ptr() + b->start(), result.ptr(), &tn));
    t.iters = tn.iters; t.error = tn.error;
    return result;
  }
}
}
}
#endif
