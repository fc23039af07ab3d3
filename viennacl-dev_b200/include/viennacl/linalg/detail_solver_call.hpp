// viennacl/linalg/detail_solver_call.hpp -- shared glue between the solver tags and the whole-solve entry points of the C-ABI.
#ifndef VIENNACL_B200_LINALG_DETAIL_SOLVER_CALL_HPP
#define VIENNACL_B200_LINALG_DETAIL_SOLVER_CALL_HPP
#include "viennacl/vector.hpp"
#include "viennacl/compressed_matrix.hpp"
#include "viennacl/sliced_ell_matrix.hpp"
#include "viennacl/ell_matrix.hpp"
#include "viennacl/hyb_matrix.hpp"
#include "viennacl/coordinate_matrix.hpp"
#include "viennacl/linalg/prod.hpp"
#include "viennacl/linalg/jacobi_precond.hpp"
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
    static ViennaCLInt trampoline(const double *x_dev, double est, void *self_)
    {
      monitor_bridge *self = static_cast<monitor_bridge*>(self_);
      viennacl::vector<NumericT> view(const_cast<NumericT*>(x_dev), CUDA_MEMORY, self->size);
      return self->fun(view, NumericT(est), self->user) ? 1 : 0;
    }
  };

  inline ViennaCLStatus call(solver_kind k, ViennaCLCUDADcsr const & A, const double *b, double *x, ViennaCLB200SolverTag *t)
  {
    ViennaCLBackend h = backend::b200::handle();
    if (k == SOLVER_CG) return ViennaCLCUDADcsr_cg(h, &A, b, x, t);
    if (k == SOLVER_BICGSTAB) return ViennaCLCUDADcsr_bicgstab(h, &A, b, x, t);
    return ViennaCLCUDADcsr_gmres(h, &A, b, x, t);
  }
  inline ViennaCLStatus call(solver_kind k, ViennaCLCUDADsell const & A, const double *b, double *x, ViennaCLB200SolverTag *t)
  {
    ViennaCLBackend h = backend::b200::handle();
    if (k == SOLVER_CG) return ViennaCLCUDADsell_cg(h, &A, b, x, t);
    if (k == SOLVER_BICGSTAB) return ViennaCLCUDADsell_bicgstab(h, &A, b, x, t);
    return ViennaCLCUDADsell_gmres(h, &A, b, x, t);
  }

  inline ViennaCLStatus call(solver_kind k, ViennaCLCUDADell const & A, const double *b, double *x, ViennaCLB200SolverTag *t)
  {
    ViennaCLBackend h = backend::b200::handle();
    if (k == SOLVER_CG) return ViennaCLCUDADell_cg(h, &A, b, x, t);
    if (k == SOLVER_BICGSTAB) return ViennaCLCUDADell_bicgstab(h, &A, b, x, t);
    return ViennaCLCUDADell_gmres(h, &A, b, x, t);
  }
  inline ViennaCLStatus call(solver_kind k, ViennaCLCUDADhyb const & A, const double *b, double *x, ViennaCLB200SolverTag *t)
  {
    ViennaCLBackend h = backend::b200::handle();
    if (k == SOLVER_CG) return ViennaCLCUDADhyb_cg(h, &A, b, x, t);
    if (k == SOLVER_BICGSTAB) return ViennaCLCUDADhyb_bicgstab(h, &A, b, x, t);
    return ViennaCLCUDADhyb_gmres(h, &A, b, x, t);
  }

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
    if (monitor) { t.monitor = &monitor_bridge<NumericT>::trampoline; t.monitor_user = &bridge; }
    else { t.monitor = NULL; t.monitor_user = NULL; }
    backend::b200::check(call(kind, A.abi(), b->ptr() + b->start(), result.ptr(), &t));
    return result;
  }
}
}
}
#endif
