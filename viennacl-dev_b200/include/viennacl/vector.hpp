// viennacl/vector.hpp -- device vector, views and the small expression vocabulary the solver drivers and the reference's
// sparse tests use (reference: vector.hpp:250-500, 949-1000, 1233-1620; detail/vector_def.hpp; vector_proxy.hpp; range.hpp).
// Elementwise work goes through the BLAS-1 entry points of the C-ABI (ViennaCLCUDADav / Davbv / Davbv_v / ...).
#ifndef VIENNACL_B200_VECTOR_HPP
#define VIENNACL_B200_VECTOR_HPP

#include <vector>
#include <cassert>
#include <algorithm>
#include <cmath>
#include <ostream>
#include "viennacl/forwards.h"
#include "viennacl/backend/mem_handle.hpp"

namespace viennacl
{

namespace tools
{
  /** @brief Rounds `to_reach` up to a multiple of `base` (tools/tools.hpp:133-137) */
  template<class IntT> IntT align_to_multiple(IntT to_reach, IntT base) { return to_reach % base == 0 ? to_reach : ((to_reach / base) + 1) * base; }
}

// ---------------------------------------------------------------------------------------------- ranges / initialisers
class range                                                                    // range.hpp
{
public:
  range() : start_(0), size_(0) {}
  range(vcl_size_t start_index, vcl_size_t stop_index) : start_(start_index), size_(stop_index - start_index) { assert(start_index <= stop_index); }
  vcl_size_t start() const { return start_; }
  vcl_size_t size() const { return size_; }
private:
  vcl_size_t start_, size_;
};

class slice                                                                    // slice.hpp
{
public:
  slice() : start_(0), stride_(1), size_(0) {}
  slice(vcl_size_t start_index, vcl_size_t stride_arg, vcl_size_t size_arg) : start_(start_index), stride_(stride_arg), size_(size_arg) {}
  vcl_size_t start() const { return start_; }
  vcl_size_t stride() const { return stride_; }
  vcl_size_t size() const { return size_; }
private:
  vcl_size_t start_, stride_, size_;
};

template<typename NumericT>
class scalar_vector                                                            // detail/vector_def.hpp:76-95
{
public:
  scalar_vector(vcl_size_t s, NumericT val, viennacl::context ctx = viennacl::context()) : size_(s), value_(val), ctx_(ctx) {}
  vcl_size_t size() const { return size_; }
  NumericT value() const { return value_; }
  viennacl::context context() const { return ctx_; }
private:
  vcl_size_t size_; NumericT value_; viennacl::context ctx_;
};

/** @brief Unit vector e_index of the given size (detail/vector_def.hpp:60-74) */
template<typename NumericT>
class unit_vector
{
public:
  unit_vector(vcl_size_t s, vcl_size_t ind, viennacl::context ctx = viennacl::context()) : size_(s), index_(ind), ctx_(ctx)
  { assert(ind < s && bool("Provided index out of range!")); }
  vcl_size_t size() const { return size_; }
  vcl_size_t index() const { return index_; }
  viennacl::context context() const { return ctx_; }
private:
  vcl_size_t size_, index_;
  viennacl::context ctx_;
};

template<typename NumericT>
class zero_vector : public scalar_vector<NumericT>
{
public:
  zero_vector(vcl_size_t s, viennacl::context ctx = viennacl::context()) : scalar_vector<NumericT>(s, 0, ctx) {}
};

// ---------------------------------------------------------------------------------------------- expressions
namespace detail
{
  /** @brief alpha*a + beta*b (+ gamma*c): every vector expression of the drivers reduces to this (avbv family) */
  template<typename NumericT>
  struct lincomb
  {
    const vector_base<NumericT> *v[3];
    NumericT c[3];
    int terms;
    lincomb() : terms(0) { v[0] = v[1] = v[2] = NULL; c[0] = c[1] = c[2] = 0; }
    void push(const vector_base<NumericT> *vec, NumericT coef)
    {
      if (terms >= 3) throw memory_exception("vector expression with more than three terms: split it");
      v[terms] = vec; c[terms] = coef; ++terms;
    }
    vcl_size_t size() const { return terms ? v[0]->size() : 0; }
  };

  template<typename NumericT>
  struct element_div_expr { const vector_base<NumericT> *num, *den; };

  /** @brief Lazy A*x (linalg/prod.hpp:350-361): evaluated on assignment by MatrixT::vec_mul (the op_executor specialisations of
   *  compressed_matrix.hpp:1231-1318 / sliced_ell_matrix.hpp:286-373, including the aliasing rule) */
  template<typename MatrixT, typename NumericT>
  struct matvec_expr { const MatrixT *A; const vector_base<NumericT> *x; };

  /** @brief v + sign * A*x, e.g. the residual `b - prod(A, x)` of examples/benchmarks/solver.cpp:117: one SpMV with beta = 1 */
  template<typename MatrixT, typename NumericT>
  struct vec_matvec_expr { const vector_base<NumericT> *v; matvec_expr<MatrixT, NumericT> Ax; NumericT sign; };
}

/** @brief Host scalar produced by a device reduction; converts implicitly like the reference's scalar_expression
 *  (scalar.hpp:82-99, 147-164). */
template<typename NumericT>
class host_scalar
{
public:
  host_scalar(NumericT v) : v_(v) {}
  operator NumericT() const { return v_; }
private:
  NumericT v_;
};

// ---------------------------------------------------------------------------------------------- element proxy / iterator
template<typename NumericT>
class entry_proxy                                                              // tools/entry_proxy.hpp:41-179: 1-element D2H/H2D
{
public:
  entry_proxy(backend::mem_handle const & h, vcl_size_t index) : h_(h), index_(index) {}
  operator NumericT() const
  {
    NumericT t;
    backend::memory_read(h_, sizeof(NumericT) * index_, sizeof(NumericT), &t);
    return t;
  }
  entry_proxy & operator=(NumericT value) { backend::memory_write(h_, sizeof(NumericT) * index_, sizeof(NumericT), &value); return *this; }
  entry_proxy & operator=(entry_proxy const & other) { NumericT t = other; return *this = t; }
  entry_proxy & operator+=(NumericT value) { NumericT t = *this; return *this = t + value; }
  entry_proxy & operator-=(NumericT value) { NumericT t = *this; return *this = t - value; }
  entry_proxy & operator*=(NumericT value) { NumericT t = *this; return *this = t * value; }
  entry_proxy & operator/=(NumericT value) { NumericT t = *this; return *this = t / value; }
private:
  backend::mem_handle h_;
  vcl_size_t index_;
};

template<typename NumericT>
class vector_iterator                                                          // vector.hpp const_vector_iterator
{
public:
  typedef vcl_ptrdiff_t difference_type;
  vector_iterator(backend::mem_handle const & h, vcl_size_t index, vcl_size_t start, vcl_size_t stride)
    : h_(h), index_(index), start_(start), stride_(stride) {}
  backend::mem_handle const & handle() const { return h_; }
  vcl_size_t index() const { return index_; }
  vcl_size_t offset() const { return start_ + index_ * stride_; }
  vcl_size_t stride() const { return stride_; }
  vector_iterator operator+(difference_type d) const { return vector_iterator(h_, vcl_size_t(difference_type(index_) + d), start_, stride_); }
  difference_type operator-(vector_iterator const & o) const { return difference_type(index_) - difference_type(o.index_); }
  bool operator==(vector_iterator const & o) const { return index_ == o.index_; }
  bool operator!=(vector_iterator const & o) const { return index_ != o.index_; }
private:
  backend::mem_handle h_;
  vcl_size_t index_, start_, stride_;
};

template<typename NumericT> void fast_copy_to_host(vector_base<NumericT> const & gpu_vec, NumericT *host);

namespace detail
{
  /** @brief y = alpha * A x + beta * y for any operator type: sparse matrix types bring vec_mul(); user-defined operators
   *  only need `apply(x, y)` computing y = A x and `size1()` (the matrix-free hook, vector.hpp:3315-3319, examples/tutorial/matrix-free.cpp) */
  template<typename MatrixT, typename NumericT>
  auto vec_mul_dispatch(MatrixT const & A, vector_base<NumericT> const & x, NumericT alpha, vector_base<NumericT> & y, NumericT beta, int)
    -> decltype(A.vec_mul(x, alpha, y, beta), void())
  { A.vec_mul(x, alpha, y, beta); }
  template<typename MatrixT, typename NumericT>
  void vec_mul_dispatch(MatrixT const & A, vector_base<NumericT> const & x, NumericT alpha, vector_base<NumericT> & y, NumericT beta, long);
}

// ---------------------------------------------------------------------------------------------- vector_base
template<typename NumericT>
class vector_base
{
public:
  typedef NumericT value_type;
  typedef NumericT cpu_value_type;
  typedef vcl_size_t size_type;
  typedef vcl_ptrdiff_t difference_type;
  typedef backend::mem_handle handle_type;
  typedef vector_iterator<NumericT> iterator;
  typedef vector_iterator<NumericT> const_iterator;

  vector_base() : size_(0), start_(0), stride_(1), internal_size_(0) {}

  /** @brief View on an existing buffer: shares the handle (vector_proxy.hpp:56-58) */
  vector_base(handle_type const & h, size_type vec_size, size_type vec_start, size_type vec_stride)
    : size_(vec_size), start_(vec_start), stride_(vec_stride), internal_size_(vec_size), elements_(h) {}

  /** @brief Creates a zero-filled vector; storage padded to a multiple of 128 entries (vector.hpp:258-267) */
  explicit vector_base(size_type vec_size, viennacl::context ctx = viennacl::context())
    : size_(vec_size), start_(0), stride_(1), internal_size_(tools::align_to_multiple<size_type>(vec_size, dense_padding_size))
  {
    check_ctx(ctx);
    if (size_ > 0) { elements_.create(sizeof(NumericT) * internal_size_); clear_all(); }
  }

  /** @brief Wraps user device memory, no ownership taken (vector.hpp:271-293) */
  explicit vector_base(NumericT *ptr_to_mem, viennacl::memory_types mem_type, size_type vec_size, vcl_size_t start = 0, size_type stride = 1)
    : size_(vec_size), start_(start), stride_(stride), internal_size_(vec_size)
  {
    if (mem_type != CUDA_MEMORY) throw memory_exception("only CUDA_MEMORY buffers can be wrapped in the B200 build");
    elements_.wrap(ptr_to_mem, sizeof(NumericT) * (start + stride * vec_size));
  }

  vector_base(vector_base const & other) : size_(0), start_(0), stride_(1), internal_size_(0) { assign_copy(other); }

  vector_base & operator=(vector_base const & other) { assign_copy(other); return *this; }

  // ---- expression assignments ----
  vector_base & operator=(detail::lincomb<NumericT> const & e) { ensure_size(e.size()); eval(e, false); return *this; }
  vector_base & operator+=(detail::lincomb<NumericT> const & e) { eval(e, true); return *this; }
  vector_base & operator-=(detail::lincomb<NumericT> e) { for (int i = 0; i < e.terms; ++i) e.c[i] = -e.c[i]; eval(e, true); return *this; }
  vector_base & operator+=(vector_base const & o) { detail::lincomb<NumericT> e; e.push(&o, NumericT(1)); eval(e, true); return *this; }
  vector_base & operator-=(vector_base const & o) { detail::lincomb<NumericT> e; e.push(&o, NumericT(-1)); eval(e, true); return *this; }
  vector_base & operator*=(NumericT s) { detail::lincomb<NumericT> e; e.push(this, s); eval(e, false); return *this; }
  vector_base & operator/=(NumericT s)
  {
    // v / s, not v * (1/s): the reference divides (vector_operations.hpp av with reciprocal flag), and so does the oracle
    backend::mem_handle den; den.create(sizeof(NumericT), &s);
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::element_div(backend::b200::handle(), int(size_), ptr(), int(start_), int(stride_),
                                                  ptr(), int(start_), int(stride_), den.ptr<NumericT>(), 0, 0));
    return *this;
  }
  vector_base & operator=(detail::element_div_expr<NumericT> const & e)
  {
    ensure_size(e.num->size());
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::element_div(backend::b200::handle(), int(size_), ptr(), int(start_), int(stride_),
                                                  e.num->ptr(), int(e.num->start()), int(e.num->stride()),
                                                  e.den->ptr(), int(e.den->start()), int(e.den->stride())));
    return *this;
  }
  template<typename MatrixT> vector_base & operator=(detail::matvec_expr<MatrixT, NumericT> const & e)
  { ensure_size(e.A->size1()); matvec(e, NumericT(1), NumericT(0)); return *this; }
  template<typename MatrixT> vector_base & operator+=(detail::matvec_expr<MatrixT, NumericT> const & e) { matvec(e, NumericT(1), NumericT(1)); return *this; }
  template<typename MatrixT> vector_base & operator-=(detail::matvec_expr<MatrixT, NumericT> const & e) { matvec(e, NumericT(-1), NumericT(1)); return *this; }

  template<typename MatrixT> vector_base & operator=(detail::vec_matvec_expr<MatrixT, NumericT> const & e)
  {
    ensure_size(e.v->size());
    if (e.Ax.x->handle() == elements_ && e.v->handle() != elements_)      // x = b - A*x: the product needs the old x
    {
      vector_base temp(size_);
      detail::vec_mul_dispatch(*e.Ax.A, *e.Ax.x, NumericT(1), temp, NumericT(0), 0);
      detail::lincomb<NumericT> l; l.push(e.v, NumericT(1)); l.push(&temp, e.sign);
      eval(l, false);
      return *this;
    }
    if (e.v->handle() != elements_ || e.v->start() != start_ || e.v->stride() != stride_) assign_copy(*e.v);
    matvec(e.Ax, e.sign, NumericT(1));
    return *this;
  }

  vector_base & operator=(scalar_vector<NumericT> const & v)
  {
    ensure_size(v.size());
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::assign(backend::b200::handle(), int(size_), ptr(), int(start_), int(stride_), v.value()));
    return *this;
  }

  vector_base & operator=(unit_vector<NumericT> const & v)
  {
    ensure_size(v.size());
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::assign(backend::b200::handle(), int(size_), ptr(), int(start_), int(stride_), NumericT(0)));
    const NumericT one = NumericT(1);
    backend::memory_write(elements_, sizeof(NumericT) * (start_ + v.index() * stride_), sizeof(NumericT), &one);
    return *this;
  }

  // ---- queries ----
  size_type size() const { return size_; }
  size_type internal_size() const { return internal_size_; }
  size_type start() const { return start_; }
  size_type stride() const { return stride_; }
  bool empty() const { return size_ == 0; }
  handle_type const & handle() const { return elements_; }
  handle_type & handle() { return elements_; }
  NumericT *ptr() const { return elements_.template ptr<NumericT>(); }
  viennacl::memory_types memory_domain() const { return elements_.get_active_handle_id(); }

  entry_proxy<NumericT> operator()(size_type index) { return entry_proxy<NumericT>(elements_, start_ + stride_ * index); }
  entry_proxy<NumericT> operator[](size_type index) { return entry_proxy<NumericT>(elements_, start_ + stride_ * index); }
  NumericT operator()(size_type index) const { return NumericT(entry_proxy<NumericT>(elements_, start_ + stride_ * index)); }
  NumericT operator[](size_type index) const { return NumericT(entry_proxy<NumericT>(elements_, start_ + stride_ * index)); }

  iterator begin() { return iterator(elements_, 0, start_, stride_); }
  iterator end() { return iterator(elements_, size_, start_, stride_); }
  const_iterator begin() const { return const_iterator(elements_, 0, start_, stride_); }
  const_iterator end() const { return const_iterator(elements_, size_, start_, stride_); }

  /** @brief Sets all entries (of this view) to zero */
  void clear()
  {
    if (size_ == 0) return;
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::assign(backend::b200::handle(), int(size_), ptr(), int(start_), int(stride_), NumericT(0)));
  }

  void resize(size_type new_size, bool preserve = true)
  {
    if (new_size == size_) return;
    vector_base fresh(new_size);
    if (preserve && size_ > 0 && new_size > 0)
    {
      size_type m = new_size < size_ ? new_size : size_;
      backend::b200::check(viennacl::backend::b200::abi<NumericT>::av(backend::b200::handle(), int(m), fresh.ptr(), 0, 1, ptr(), int(start_), int(stride_), NumericT(1)));
    }
    swap(fresh);
  }

  void swap(vector_base & other)
  {
    std::swap(size_, other.size_); std::swap(start_, other.start_); std::swap(stride_, other.stride_);
    std::swap(internal_size_, other.internal_size_); elements_.swap(other.elements_);
  }

protected:
  void check_ctx(viennacl::context const & ctx) const
  {
    if (ctx.memory_type() != CUDA_MEMORY) throw memory_exception("only CUDA_MEMORY is available in the B200 build (no host/OpenCL backend)");
  }
  void clear_all() { backend::b200::check(ViennaCLCUDAMemSet(backend::b200::handle(), elements_.get(), 0, sizeof(NumericT) * internal_size_)); }
  void ensure_size(size_type s)
  {
    if (size_ == 0 && s > 0)
    {
      size_ = s; start_ = 0; stride_ = 1; internal_size_ = tools::align_to_multiple<size_type>(s, dense_padding_size);
      elements_.create(sizeof(NumericT) * internal_size_); clear_all();
    }
    assert(size_ == s && bool("Incompatible vector sizes!"));
  }
  void assign_copy(vector_base const & other)
  {
    if (&other == this || other.size() == 0) return;
    ensure_size(other.size());
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::av(backend::b200::handle(), int(size_), ptr(), int(start_), int(stride_),
                                         other.ptr(), int(other.start()), int(other.stride()), NumericT(1)));
  }
  template<typename MatrixT>
  void matvec(detail::matvec_expr<MatrixT, NumericT> const & e, NumericT alpha, NumericT beta)
  {
    assert(e.A->size1() == size_ && bool("Size check failed for matrix-vector product"));
    if (e.x->handle() == elements_)         // x = A*x and friends: go through a temporary (compressed_matrix.hpp:1237-1242)
    {
      vector_base temp(size_);
      detail::vec_mul_dispatch(*e.A, *e.x, NumericT(1), temp, NumericT(0), 0);
      detail::lincomb<NumericT> l; l.push(&temp, alpha);
      if (beta != NumericT(0)) eval(l, true); else eval(l, false);
    }
    else
      detail::vec_mul_dispatch(*e.A, *e.x, alpha, *this, beta, 0);
  }
  // Does operand o share memory with this vector?  0: no; 1: the very same view (element i of o IS element i of *this: in-place
  // element-wise evaluation is safe); 2: another view of the same buffer (overlap possible: evaluate through a temporary)
  int alias_kind(vector_base const * o) const
  {
    if (!(o->handle() == elements_)) return 0;
    return (o->start() == start_ && o->stride() == stride_) ? 1 : 2;
  }
  void eval(detail::lincomb<NumericT> const & e_in, bool accumulate)
  {
    if (e_in.terms == 0) return;
    assert(e_in.size() == size_ && bool("Incompatible vector sizes!"));
    detail::lincomb<NumericT> e = e_in;
    // Aliasing (the reference evaluates such expressions through temporaries, vector.hpp op_executor specialisations): a
    // three-term expression takes two passes, and the operand of the SECOND pass must not be the destination itself --
    // x = a + b + x would otherwise add the already overwritten x.  Reorder so that aliasing operands are read by the first
    // pass (in place, element by element); different views of the destination's buffer, or three aliasing operands, go
    // through a temporary.
    {
      int kinds[3] = {0, 0, 0}, n_alias = 0; bool other_view = false;
      for (int k = 0; k < e.terms; ++k) { kinds[k] = alias_kind(e.v[k]); n_alias += kinds[k] != 0; other_view |= kinds[k] == 2; }
      if (other_view || (e.terms == 3 && n_alias == 3))
      {
        vector_base temp(size_);
        temp.eval(e, false);
        detail::lincomb<NumericT> l; l.push(&temp, NumericT(1));
        eval(l, accumulate);
        return;
      }
      if (e.terms == 3 && kinds[2] != 0)
      {
        const int swap_with = kinds[0] == 0 ? 0 : 1;          // a non-aliasing operand of the first pass
        std::swap(e.v[2], e.v[swap_with]); std::swap(e.c[2], e.c[swap_with]);
      }
    }
    ViennaCLBackend b = backend::b200::handle();
    const vector_base *a = e.v[0], *c = e.terms > 1 ? e.v[1] : e.v[0];
    const NumericT ca = e.c[0], cc = e.terms > 1 ? e.c[1] : NumericT(0);
    if (!accumulate && e.terms == 1)
      backend::b200::check(viennacl::backend::b200::abi<NumericT>::av(b, int(size_), ptr(), int(start_), int(stride_), a->ptr(), int(a->start()), int(a->stride()), ca));
    else if (!accumulate)
      backend::b200::check(viennacl::backend::b200::abi<NumericT>::avbv(b, int(size_), ptr(), int(start_), int(stride_), a->ptr(), int(a->start()), int(a->stride()), ca,
                                             c->ptr(), int(c->start()), int(c->stride()), cc));
    else
      backend::b200::check(viennacl::backend::b200::abi<NumericT>::avbv_v(b, int(size_), ptr(), int(start_), int(stride_), a->ptr(), int(a->start()), int(a->stride()), ca,
                                               c->ptr(), int(c->start()), int(c->stride()), cc));
    if (e.terms == 3)
    {
      const vector_base *d = e.v[2];
      backend::b200::check(viennacl::backend::b200::abi<NumericT>::avbv_v(b, int(size_), ptr(), int(start_), int(stride_), d->ptr(), int(d->start()), int(d->stride()), e.c[2],
                                               d->ptr(), int(d->start()), int(d->stride()), NumericT(0)));
    }
  }

  size_type size_, start_, stride_, internal_size_;
  handle_type elements_;
};

// ---------------------------------------------------------------------------------------------- vector
template<typename NumericT, unsigned int AlignmentV>
class vector : public vector_base<NumericT>
{
  typedef vector_base<NumericT> base_type;
public:
  typedef typename base_type::size_type size_type;
  typedef typename base_type::difference_type difference_type;

  vector() : base_type() {}
  explicit vector(size_type vec_size) : base_type(vec_size) {}
  explicit vector(size_type vec_size, viennacl::context ctx) : base_type(vec_size, ctx) {}
  explicit vector(NumericT *ptr_to_mem, viennacl::memory_types mem_type, size_type vec_size, size_type start = 0, size_type stride = 1)
    : base_type(ptr_to_mem, mem_type, vec_size, start, stride) {}
  vector(base_type const & v) : base_type(v) {}
  vector(vector const & v) : base_type(v) {}
  vector(detail::lincomb<NumericT> const & e) : base_type() { base_type::operator=(e); }
  vector(scalar_vector<NumericT> const & v) : base_type(v.size(), v.context()) { if (v.value() != NumericT(0)) base_type::operator=(v); }
  vector(zero_vector<NumericT> const & v) : base_type(v.size(), v.context()) {}
  vector(unit_vector<NumericT> const & v) : base_type(v.size(), v.context()) { base_type::operator=(v); }
  template<typename MatrixT> vector(detail::matvec_expr<MatrixT, NumericT> const & e) : base_type() { base_type::operator=(e); }
  template<typename MatrixT> vector(detail::vec_matvec_expr<MatrixT, NumericT> const & e) : base_type() { base_type::operator=(e); }

  vector & operator=(vector const & o) { base_type::operator=(static_cast<base_type const &>(o)); return *this; }
  using base_type::operator=;

  void swap(vector & other) { base_type::swap(other); }
};

// y = beta * y (beta == 0: y = 0 without reading it): the product with a matrix that holds no entries, e.g. after clear()
// (tests/src/sparse.cpp:988-1060 "products after clear()")
namespace detail
{
  template<typename NumericT>
  void scale_by_beta(vector_base<NumericT> & y, NumericT beta)
  {
    if (y.size() == 0 || beta == NumericT(1)) return;
    if (beta == NumericT(0)) { y.clear(); return; }
    backend::b200::check(viennacl::backend::b200::abi<NumericT>::av(backend::b200::handle(), int(y.size()), y.ptr(), int(y.start()), int(y.stride()),
                                                                  y.ptr(), int(y.start()), int(y.stride()), beta));
  }
}

// ---------------------------------------------------------------------------------------------- views
template<typename VectorType>
class vector_range : public vector_base<typename VectorType::value_type>       // vector_proxy.hpp:40-100
{
  typedef vector_base<typename VectorType::value_type> base_type;
public:
  vector_range(VectorType const & v, range const & r) : base_type(v.handle(), r.size(), v.start() + v.stride() * r.start(), v.stride()) {}
  using base_type::operator=;
};

template<typename VectorType>
class vector_slice : public vector_base<typename VectorType::value_type>       // vector_proxy.hpp:215-330
{
  typedef vector_base<typename VectorType::value_type> base_type;
public:
  vector_slice(VectorType const & v, slice const & s) : base_type(v.handle(), s.size(), v.start() + v.stride() * s.start(), v.stride() * s.stride()) {}
  using base_type::operator=;
};

template<typename VectorType> vector_range<VectorType> project(VectorType const & vec, viennacl::range const & r1) { return vector_range<VectorType>(vec, r1); }
template<typename VectorType> vector_slice<VectorType> project(VectorType const & vec, viennacl::slice const & s1) { return vector_slice<VectorType>(vec, s1); }

// ---------------------------------------------------------------------------------------------- operators -> lincomb
template<typename T> detail::lincomb<T> operator*(T a, vector_base<T> const & v) { detail::lincomb<T> e; e.push(&v, a); return e; }
template<typename T> detail::lincomb<T> operator*(vector_base<T> const & v, T a) { detail::lincomb<T> e; e.push(&v, a); return e; }
template<typename T> detail::lincomb<T> operator-(vector_base<T> const & v) { detail::lincomb<T> e; e.push(&v, T(-1)); return e; }
template<typename T> detail::lincomb<T> operator*(T a, detail::lincomb<T> e) { for (int i = 0; i < e.terms; ++i) e.c[i] *= a; return e; }
template<typename T> detail::lincomb<T> operator*(detail::lincomb<T> e, T a) { for (int i = 0; i < e.terms; ++i) e.c[i] *= a; return e; }
template<typename T> detail::lincomb<T> operator+(vector_base<T> const & a, vector_base<T> const & b) { detail::lincomb<T> e; e.push(&a, T(1)); e.push(&b, T(1)); return e; }
template<typename T> detail::lincomb<T> operator-(vector_base<T> const & a, vector_base<T> const & b) { detail::lincomb<T> e; e.push(&a, T(1)); e.push(&b, T(-1)); return e; }
template<typename T> detail::lincomb<T> operator+(vector_base<T> const & a, detail::lincomb<T> const & b) { detail::lincomb<T> e; e.push(&a, T(1)); for (int i = 0; i < b.terms; ++i) e.push(b.v[i], b.c[i]); return e; }
template<typename T> detail::lincomb<T> operator-(vector_base<T> const & a, detail::lincomb<T> const & b) { detail::lincomb<T> e; e.push(&a, T(1)); for (int i = 0; i < b.terms; ++i) e.push(b.v[i], -b.c[i]); return e; }
template<typename T> detail::lincomb<T> operator+(detail::lincomb<T> e, vector_base<T> const & b) { e.push(&b, T(1)); return e; }
template<typename T> detail::lincomb<T> operator-(detail::lincomb<T> e, vector_base<T> const & b) { e.push(&b, T(-1)); return e; }
template<typename T> detail::lincomb<T> operator+(detail::lincomb<T> e, detail::lincomb<T> const & b) { for (int i = 0; i < b.terms; ++i) e.push(b.v[i], b.c[i]); return e; }
template<typename T> detail::lincomb<T> operator-(detail::lincomb<T> e, detail::lincomb<T> const & b) { for (int i = 0; i < b.terms; ++i) e.push(b.v[i], -b.c[i]); return e; }

template<typename M, typename T> detail::vec_matvec_expr<M, T> operator+(vector_base<T> const & v, detail::matvec_expr<M, T> const & e)
{ detail::vec_matvec_expr<M, T> r = {&v, e, T(1)}; return r; }
template<typename M, typename T> detail::vec_matvec_expr<M, T> operator-(vector_base<T> const & v, detail::matvec_expr<M, T> const & e)
{ detail::vec_matvec_expr<M, T> r = {&v, e, T(-1)}; return r; }

/** @brief Prints "[n](v0,v1,...)" like the reference (vector.hpp:1820-1840) */
template<typename T>
std::ostream & operator<<(std::ostream & os, vector_base<T> const & val)
{
  std::vector<T> tmp(val.size());
  if (!tmp.empty()) fast_copy_to_host(val, &tmp[0]);
  os << "[" << val.size() << "](";
  for (vcl_size_t i = 0; i < tmp.size(); ++i) { if (i > 0) os << ","; os << tmp[i]; }
  os << ")";
  return os;
}

namespace detail
{
  template<typename MatrixT, typename NumericT>
  void vec_mul_dispatch(MatrixT const & A, vector_base<NumericT> const & x, NumericT alpha, vector_base<NumericT> & y, NumericT beta, long)
  {
    if (alpha == NumericT(1) && beta == NumericT(0)) { A.apply(x, y); return; }
    vector<NumericT> t(y.size());
    A.apply(x, t);
    if (beta == NumericT(0)) y = alpha * t;
    else y = alpha * t + beta * y;
  }
}

namespace linalg
{
  template<typename T> viennacl::detail::element_div_expr<T> element_div(vector_base<T> const & a, vector_base<T> const & b)
  { viennacl::detail::element_div_expr<T> e = {&a, &b}; return e; }
}

// ---------------------------------------------------------------------------------------------- host <-> device copies
/** @brief fast_copy host -> device: contiguous host range into a (possibly strided) device range (vector.hpp:1233-1312) */
template<typename NumericT, typename CPUIt>
void fast_copy(CPUIt const & cpu_begin, CPUIt const & cpu_end, vector_iterator<NumericT> gpu_begin)
{
  vcl_size_t n = vcl_size_t(cpu_end - cpu_begin);
  if (n == 0) return;
  if (gpu_begin.stride() == 1)
    backend::memory_write(const_cast<backend::mem_handle &>(gpu_begin.handle()), sizeof(NumericT) * gpu_begin.offset(), sizeof(NumericT) * n, &(*cpu_begin));
  else
  {
    vector<NumericT> tmp(n);
    backend::memory_write(tmp.handle(), 0, sizeof(NumericT) * n, &(*cpu_begin));
    vector_base<NumericT> dst(gpu_begin.handle(), n, gpu_begin.offset(), gpu_begin.stride());
    dst = tmp;
  }
}

/** @brief fast_copy device -> host (blocking, like backend/cuda.hpp:183-200) */
template<typename NumericT, typename CPUIt>
void fast_copy(vector_iterator<NumericT> const & gpu_begin, vector_iterator<NumericT> const & gpu_end, CPUIt cpu_begin)
{
  vcl_size_t n = vcl_size_t(gpu_end - gpu_begin);
  if (n == 0) return;
  if (gpu_begin.stride() == 1)
    backend::memory_read(gpu_begin.handle(), sizeof(NumericT) * gpu_begin.offset(), sizeof(NumericT) * n, &(*cpu_begin));
  else
  {
    vector<NumericT> tmp(n);
    vector_base<NumericT> src(gpu_begin.handle(), n, gpu_begin.offset(), gpu_begin.stride());
    static_cast<vector_base<NumericT> &>(tmp) = src;
    backend::memory_read(tmp.handle(), 0, sizeof(NumericT) * n, &(*cpu_begin));
  }
}

template<typename NumericT>
void fast_copy_to_host(vector_base<NumericT> const & gpu_vec, NumericT *host) { fast_copy(gpu_vec.begin(), gpu_vec.end(), host); }

template<typename NumericT, typename CPUVectorT>
void fast_copy(vector_base<NumericT> const & gpu_vec, CPUVectorT & cpu_vec) { fast_copy(gpu_vec.begin(), gpu_vec.end(), cpu_vec.begin()); }
template<typename NumericT, typename CPUVectorT>
void fast_copy(CPUVectorT const & cpu_vec, vector_base<NumericT> & gpu_vec) { fast_copy(cpu_vec.begin(), cpu_vec.end(), gpu_vec.begin()); }

/** @brief copy(): same as fast_copy for contiguous STL containers (vector.hpp:1333-1620) */
template<typename NumericT, typename CPUIt>
void copy(CPUIt const & cpu_begin, CPUIt const & cpu_end, vector_iterator<NumericT> gpu_begin)
{
  std::vector<NumericT> tmp(cpu_begin, cpu_end);
  if (!tmp.empty()) fast_copy(tmp.begin(), tmp.end(), gpu_begin);
}
template<typename NumericT, typename CPUIt>
void copy(vector_iterator<NumericT> const & gpu_begin, vector_iterator<NumericT> const & gpu_end, CPUIt cpu_begin)
{
  std::vector<NumericT> tmp(vcl_size_t(gpu_end - gpu_begin));
  if (tmp.empty()) return;
  fast_copy(gpu_begin, gpu_end, tmp.begin());
  for (vcl_size_t i = 0; i < tmp.size(); ++i, ++cpu_begin) *cpu_begin = tmp[i];
}
template<typename NumericT, typename CPUVectorT>
void copy(CPUVectorT const & cpu_vec, vector_base<NumericT> & gpu_vec) { viennacl::copy(cpu_vec.begin(), cpu_vec.end(), gpu_vec.begin()); }
template<typename NumericT, typename CPUVectorT>
void copy(vector_base<NumericT> const & gpu_vec, CPUVectorT & cpu_vec) { viennacl::copy(gpu_vec.begin(), gpu_vec.end(), cpu_vec.begin()); }

/** @brief async_copy (vector.hpp:1276-1312, :1400-1440): the transfer is enqueued on the backend's stream and NOT waited for; the
 *  host range must stay valid (and, device -> host, unread) until backend::finish().  Contiguous ranges only. */
template<typename NumericT, typename CPUIt>
void async_copy(CPUIt const & cpu_begin, CPUIt const & cpu_end, vector_iterator<NumericT> gpu_begin)
{
  vcl_size_t n = vcl_size_t(cpu_end - cpu_begin);
  if (n == 0) return;
  assert(gpu_begin.stride() == 1 && bool("async_copy needs a contiguous device range"));
  backend::memory_write(const_cast<backend::mem_handle &>(gpu_begin.handle()), sizeof(NumericT) * gpu_begin.offset(), sizeof(NumericT) * n, &(*cpu_begin), true);
}
template<typename NumericT, typename CPUIt>
void async_copy(vector_iterator<NumericT> const & gpu_begin, vector_iterator<NumericT> const & gpu_end, CPUIt cpu_begin)
{
  vcl_size_t n = vcl_size_t(gpu_end - gpu_begin);
  if (n == 0) return;
  assert(gpu_begin.stride() == 1 && bool("async_copy needs a contiguous device range"));
  backend::memory_read(gpu_begin.handle(), sizeof(NumericT) * gpu_begin.offset(), sizeof(NumericT) * n, &(*cpu_begin), true);
}
template<typename NumericT, typename CPUVectorT>
void async_copy(CPUVectorT const & cpu_vec, vector_base<NumericT> & gpu_vec) { viennacl::async_copy(cpu_vec.begin(), cpu_vec.end(), gpu_vec.begin()); }
template<typename NumericT, typename CPUVectorT>
void async_copy(vector_base<NumericT> const & gpu_vec, CPUVectorT & cpu_vec) { viennacl::async_copy(gpu_vec.begin(), gpu_vec.end(), cpu_vec.begin()); }

namespace traits
{
  template<typename T> vcl_size_t size(vector_base<T> const & v) { return v.size(); }
  template<typename T> vcl_size_t start(vector_base<T> const & v) { return v.start(); }
  template<typename T> vcl_size_t stride(vector_base<T> const & v) { return v.stride(); }
  template<typename T> backend::mem_handle const & handle(vector_base<T> const & v) { return v.handle(); }
  template<typename T> void clear(vector_base<T> & v) { v.clear(); }
  template<typename T> viennacl::context context(T const &) { return viennacl::context(CUDA_MEMORY); }
}

} // namespace viennacl
#endif
