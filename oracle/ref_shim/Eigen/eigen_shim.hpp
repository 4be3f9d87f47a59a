// TEST INFRASTRUCTURE — not part of the product.
//
// A minimal, eager stand-in for the parts of the Eigen 3.4 API that plainmp's hot-path sources use
// (cpp/plainmp/{kinematics,collision,constraints}/..., ompl/motion_validator.cpp), so that those
// sources can be compiled UNMODIFIED from /root/reference into oracle/_ref/ on a machine that has
// no Eigen (oracle/Makefile.ref).  Eigen itself is header-only third-party code that is absent from
// this image and from the reference tree; what is restated here is its published arithmetic:
//   * Quaternion product, _transformVector, toRotationMatrix, inverse  (Eigen/src/Geometry/Quaternion.h)
//   * fixed-size 3x3 inverse by cofactors                              (Eigen/src/LU/InverseImpl.h)
//   * isApprox                                                          (Eigen/src/Core/Fuzzy.h)
//   * small fixed-size reductions are evaluated with Eigen's unrolled pairwise split
//     (redux_novec_unroller: f(first half, second half)), i.e. a0 + (a1 + a2) for three terms,
//     and small products coefficient-wise as (lhs.row(i).cwiseProduct(rhs.col(j))).sum().
// The reference builds with EIGEN_DONT_VECTORIZE (CMakeLists.txt:10), so scalar evaluation order
// is what upstream executes too.
//
// No expression templates: every operator returns a plain Matrix by value.  Only the members the
// reference sources touch exist.
#pragma once
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <type_traits>
#include <vector>

namespace Eigen {

constexpr int Dynamic = -1;
using Index = std::ptrdiff_t;

template <class S, int R, int C>
class Matrix;
template <class S, int R, int C>
struct View;
template <class S, int R, int C>
struct ArrayOf;

namespace detail {
template <class S, int R, int C, bool Fixed = (R != Dynamic && C != Dynamic)>
struct Storage;
template <class S, int R, int C>
struct Storage<S, R, C, true> {
  S d[R * C];
  Storage() {}
  Storage(Index, Index) {}
  Index rows() const { return R; }
  Index cols() const { return C; }
  void resize(Index, Index) {}
  S* data() { return d; }
  const S* data() const { return d; }
};
template <class S, int R, int C>
struct Storage<S, R, C, false> {
  std::vector<S> d;
  Index r = (R == Dynamic ? 0 : R), c = (C == Dynamic ? 0 : C);
  Storage() {}
  Storage(Index rr, Index cc) : d(size_t(rr * cc)), r(rr), c(cc) {}
  Index rows() const { return r; }
  Index cols() const { return c; }
  void resize(Index rr, Index cc) {
    r = rr;
    c = cc;
    d.resize(size_t(rr * cc));
  }
  S* data() { return d.data(); }
  const S* data() const { return d.data(); }
};
constexpr int pick(int a, int b) { return a != Dynamic ? a : b; }

// Eigen's unrolled reduction order for small fixed sizes: f(reduce(first half), reduce(rest)).
template <class F, class G>
auto redux_tree(Index start, Index len, const G& get, const F& f) -> decltype(get(Index(0))) {
  if (len == 1) return get(start);
  const Index half = len / 2;
  return f(redux_tree(start, half, get, f), redux_tree(start + half, len - half, get, f));
}
// Fixed sizes are completely unrolled by Eigen (pairwise split); dynamic sizes run a linear loop.
template <bool Fixed, class F, class G>
auto redux(Index n, const G& get, const F& f) -> decltype(get(Index(0))) {
  if (Fixed) return redux_tree(0, n, get, f);
  auto acc = get(0);
  for (Index i = 1; i < n; i++) acc = f(acc, get(i));
  return acc;
}
}  // namespace detail

template <class D>
struct Traits;

template <class M>
struct ColwiseProxy;
template <class M>
struct RowwiseProxy;
template <class M>
struct CommaInit;

// ---------------------------------------------------------------------------------------------
template <class Derived>
struct DenseBase {
  using Scalar = typename Traits<Derived>::Scalar;
  static constexpr int RowsAtCompileTime = Traits<Derived>::Rows;
  static constexpr int ColsAtCompileTime = Traits<Derived>::Cols;
  static constexpr bool IsFixed = RowsAtCompileTime != Dynamic && ColsAtCompileTime != Dynamic;
  using Plain = Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime>;
  using S = Scalar;

  const Derived& derived() const { return *static_cast<const Derived*>(this); }
  Derived& derived() { return *static_cast<Derived*>(this); }
  Index rows() const { return derived().rows_(); }
  Index cols() const { return derived().cols_(); }
  Index size() const { return rows() * cols(); }
  S coeff(Index i, Index j) const { return derived().at(i, j); }
  S& coeffRef(Index i, Index j) { return derived().at(i, j); }
  // linear (vector) access
  S lin(Index k) const { return cols() == 1 ? coeff(k, 0) : (rows() == 1 ? coeff(0, k) : coeff(k % rows(), k / rows())); }
  S& linRef(Index k) { return cols() == 1 ? coeffRef(k, 0) : (rows() == 1 ? coeffRef(0, k) : coeffRef(k % rows(), k / rows())); }

  S operator()(Index i, Index j) const { return coeff(i, j); }
  S& operator()(Index i, Index j) { return coeffRef(i, j); }
  S operator()(Index k) const { return lin(k); }
  S& operator()(Index k) { return linRef(k); }
  S operator[](Index k) const { return lin(k); }
  S& operator[](Index k) { return linRef(k); }
  S x() const { return lin(0); }
  S y() const { return lin(1); }
  S z() const { return lin(2); }
  S w() const { return lin(3); }
  S& x() { return linRef(0); }
  S& y() { return linRef(1); }
  S& z() { return linRef(2); }
  S& w() { return linRef(3); }

  Plain eval() const {
    Plain r(rows(), cols());
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) r.at(i, j) = coeff(i, j);
    return r;
  }

  // ---- reductions
  S sum() const {
    return detail::redux<IsFixed>(size(), [&](Index k) { return lin(k); }, [](S a, S b) { return a + b; });
  }
  S squaredNorm() const {
    return detail::redux<IsFixed>(size(), [&](Index k) { return lin(k) * lin(k); }, [](S a, S b) { return a + b; });
  }
  S norm() const { return std::sqrt(squaredNorm()); }
  S maxCoeff() const {
    return detail::redux<IsFixed>(size(), [&](Index k) { return lin(k); }, [](S a, S b) { return a < b ? b : a; });
  }
  S minCoeff() const {
    return detail::redux<IsFixed>(size(), [&](Index k) { return lin(k); }, [](S a, S b) { return b < a ? b : a; });
  }
  template <class O>
  S dot(const DenseBase<O>& o) const {
    constexpr bool fx = IsFixed || DenseBase<O>::IsFixed;
    return detail::redux<fx>(size(), [&](Index k) { return lin(k) * o.lin(k); }, [](S a, S b) { return a + b; });
  }
  template <class O>
  Matrix<S, 3, 1> cross(const DenseBase<O>& o) const {
    return Matrix<S, 3, 1>(lin(1) * o.lin(2) - lin(2) * o.lin(1), lin(2) * o.lin(0) - lin(0) * o.lin(2),
                           lin(0) * o.lin(1) - lin(1) * o.lin(0));
  }
  template <class O>
  bool isApprox(const DenseBase<O>& o, S prec = S(1e-12)) const {
    S diff = (*this - o).squaredNorm();
    return diff <= prec * prec * std::min(squaredNorm(), o.squaredNorm());
  }

  // ---- coefficient-wise
  template <class F>
  Plain unary(const F& f) const {
    Plain r(rows(), cols());
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) r.at(i, j) = f(coeff(i, j));
    return r;
  }
  Plain cwiseAbs() const {
    return unary([](S a) { return std::abs(a); });
  }
  Plain cwiseMin(S v) const {
    return unary([v](S a) { return std::min(a, v); });
  }
  Plain cwiseMax(S v) const {
    return unary([v](S a) { return std::max(a, v); });
  }
  template <class O>
  Plain cwiseMin(const DenseBase<O>& o) const {
    Plain r(rows(), cols());
    for (Index k = 0; k < size(); k++) r.linRef(k) = std::min(lin(k), o.lin(k));
    return r;
  }
  template <class O>
  Plain cwiseProduct(const DenseBase<O>& o) const {
    Plain r(rows(), cols());
    for (Index k = 0; k < size(); k++) r.linRef(k) = lin(k) * o.lin(k);
    return r;
  }
  Plain operator-() const {
    return unary([](S a) { return -a; });
  }
  Matrix<S, ColsAtCompileTime, RowsAtCompileTime> transpose() const {
    Matrix<S, ColsAtCompileTime, RowsAtCompileTime> r(cols(), rows());
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) r.at(j, i) = coeff(i, j);
    return r;
  }
  template <class T>
  Matrix<T, RowsAtCompileTime, ColsAtCompileTime> cast() const {
    Matrix<T, RowsAtCompileTime, ColsAtCompileTime> r(rows(), cols());
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) r.at(i, j) = T(coeff(i, j));
    return r;
  }
  // fixed 3x3 only (Pose::rot_inv_): cofactors / determinant, Eigen/src/LU/InverseImpl.h
  Plain inverse() const {
    static_assert(RowsAtCompileTime == 3 && ColsAtCompileTime == 3, "shim: only 3x3 inverse");
    auto cof = [&](int i, int j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return coeff(i1, j1) * coeff(i2, j2) - coeff(i1, j2) * coeff(i2, j1);
    };
    const S c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    // (cofactors_col0.cwiseProduct(matrix.col(0))).sum(): unrolled pairwise
    const S det = c00 * coeff(0, 0) + (c10 * coeff(1, 0) + c20 * coeff(2, 0));
    const S invdet = S(1) / det;
    Plain r;
    r.at(0, 0) = c00 * invdet;
    r.at(0, 1) = c10 * invdet;
    r.at(0, 2) = c20 * invdet;
    r.at(1, 0) = cof(0, 1) * invdet;
    r.at(1, 1) = cof(1, 1) * invdet;
    r.at(1, 2) = cof(2, 1) * invdet;
    r.at(2, 0) = cof(0, 2) * invdet;
    r.at(2, 1) = cof(1, 2) * invdet;
    r.at(2, 2) = cof(2, 2) * invdet;
    return r;
  }

  // ---- in-place
  template <class O>
  Derived& assign_from(const DenseBase<O>& o) {
    if (rows() == o.rows() && cols() == o.cols()) {
      for (Index j = 0; j < cols(); j++)
        for (Index i = 0; i < rows(); i++) coeffRef(i, j) = S(o.coeff(i, j));
    } else {  // Eigen lets a row vector be assigned to a column vector and vice versa
      assert(size() == o.size() && (rows() == 1 || cols() == 1) && (o.rows() == 1 || o.cols() == 1));
      for (Index k = 0; k < size(); k++) linRef(k) = S(o.lin(k));
    }
    return derived();
  }
  Derived& setConstant(S v) {
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) coeffRef(i, j) = v;
    return derived();
  }
  Derived& setZero() { return setConstant(S(0)); }
  template <class O>
  Derived& operator+=(const DenseBase<O>& o) {
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) coeffRef(i, j) += o.coeff(i, j);
    return derived();
  }
  template <class O>
  Derived& operator-=(const DenseBase<O>& o) {
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) coeffRef(i, j) -= o.coeff(i, j);
    return derived();
  }
  template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>
  Derived& operator*=(T s) {
    for (Index k = 0; k < size(); k++) linRef(k) *= S(s);
    return derived();
  }
  template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>
  Derived& operator/=(T s) {
    for (Index k = 0; k < size(); k++) linRef(k) /= S(s);
    return derived();
  }

  // ---- blocks
  View<S, RowsAtCompileTime, 1> col(Index j) { return {&coeffRef(0, j), rows(), 1, derived().rstride_(), derived().cstride_()}; }
  View<S, RowsAtCompileTime, 1> col(Index j) const {
    return const_cast<DenseBase*>(this)->col(j);
  }
  View<S, 1, ColsAtCompileTime> row(Index i) { return {&coeffRef(i, 0), 1, cols(), derived().rstride_(), derived().cstride_()}; }
  View<S, 1, ColsAtCompileTime> row(Index i) const { return const_cast<DenseBase*>(this)->row(i); }
  template <int BR, int BC>
  View<S, BR, BC> block(Index i, Index j) {
    return {&coeffRef(i, j), BR, BC, derived().rstride_(), derived().cstride_()};
  }
  // run-time sized block / segment (composite_constraint.hpp:57-58)
  View<S, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) {
    return {&coeffRef(i, j), r, c, derived().rstride_(), derived().cstride_()};
  }
  View<S, Dynamic, 1> segment(Index i, Index n) { return {&linRef(i), n, 1, cols() == 1 ? derived().rstride_() : derived().cstride_(), 0}; }
  View<S, Dynamic, 1> head(Index n) { return {&linRef(0), n, 1, cols() == 1 ? derived().rstride_() : derived().cstride_(), 0}; }
  View<S, Dynamic, 1> head(Index n) const { return const_cast<DenseBase*>(this)->head(n); }
  ArrayOf<S, RowsAtCompileTime, ColsAtCompileTime> array() {
    return {View<S, RowsAtCompileTime, ColsAtCompileTime>{&coeffRef(0, 0), rows(), cols(), derived().rstride_(), derived().cstride_()}};
  }
  ArrayOf<S, RowsAtCompileTime, ColsAtCompileTime> array() const { return const_cast<DenseBase*>(this)->array(); }
  ColwiseProxy<Derived> colwise() const { return {derived()}; }
  RowwiseProxy<Derived> rowwise() const { return {derived()}; }

  CommaInit<Derived> operator<<(S v) {
    CommaInit<Derived> ci{derived(), 0};
    ci, v;
    return ci;
  }
  template <class O>
  CommaInit<Derived> operator<<(const DenseBase<O>& v) {
    CommaInit<Derived> ci{derived(), 0};
    ci, v;
    return ci;
  }
};

// m << a, b, c ...: row-major fill (scalars, or vectors appended coefficient by coefficient)
template <class M>
struct CommaInit {
  M& m;
  Index k;
  typename Traits<M>::Scalar& slot() {
    const Index r = k / m.cols(), c = k % m.cols();
    k++;
    return m.coeffRef(r, c);
  }
  CommaInit& operator,(typename Traits<M>::Scalar v) {
    slot() = v;
    return *this;
  }
  template <class O>
  CommaInit& operator,(const DenseBase<O>& v) {
    for (Index i = 0; i < v.size(); i++) slot() = v.lin(i);
    return *this;
  }
};

// ---------------------------------------------------------------------------------------------
template <class S_, int R, int C>
struct Traits<Matrix<S_, R, C>> {
  using Scalar = S_;
  static constexpr int Rows = R, Cols = C;
};
template <class S_, int R, int C>
struct Traits<View<S_, R, C>> {
  using Scalar = S_;
  static constexpr int Rows = R, Cols = C;
};

template <class S, int R, int C>
class Matrix : public DenseBase<Matrix<S, R, C>> {
  detail::Storage<S, R, C> st_;

 public:
  using Base = DenseBase<Matrix<S, R, C>>;
  using Scalar = S;
  Matrix() {}
  // size constructors (dynamic types) / 2-vector value constructor
  template <class T, class = std::enable_if_t<std::is_integral<T>::value && (R == Dynamic || C == Dynamic)>>
  explicit Matrix(T n) : st_(R == 1 ? 1 : Index(n), R == 1 ? Index(n) : 1) {}
  template <class T0, class T1,
            class = std::enable_if_t<std::is_arithmetic<T0>::value && std::is_arithmetic<T1>::value>>
  Matrix(T0 a, T1 b) : st_(Index(R * C == 2 ? 2 : Index(a)), Index(R * C == 2 ? 1 : Index(b))) {
    if (R * C == 2) {
      st_.data()[0] = S(a);
      st_.data()[1] = S(b);
    }
  }
  Matrix(S a, S b, S c) {
    static_assert(R * C == 3, "3-vector ctor");
    st_.d[0] = a;
    st_.d[1] = b;
    st_.d[2] = c;
  }
  Matrix(S a, S b, S c, S d) {
    static_assert(R * C == 4, "4-vector ctor");
    st_.d[0] = a;
    st_.d[1] = b;
    st_.d[2] = c;
    st_.d[3] = d;
  }
  template <class O>
  Matrix(const DenseBase<O>& o) : st_(rows_for(o), cols_for(o)) {
    this->assign_from(o);
  }
  Matrix(const ArrayOf<S, R, C>& a);
  template <int R2, int C2>
  Matrix(const ArrayOf<S, R2, C2>& a);
  Matrix(const Matrix&) = default;
  Matrix(Matrix&&) = default;
  Matrix& operator=(const Matrix&) = default;
  Matrix& operator=(Matrix&&) = default;
  template <class O>
  Matrix& operator=(const DenseBase<O>& o) {
    st_.resize(rows_for(o), cols_for(o));
    return this->assign_from(o);
  }
  template <int R2, int C2>
  Matrix& operator=(const ArrayOf<S, R2, C2>& a);

  template <class O>
  static Index rows_for(const DenseBase<O>& o) {
    if (R != Dynamic) return R;
    return C == 1 ? o.size() : o.rows();
  }
  template <class O>
  static Index cols_for(const DenseBase<O>& o) {
    if (C != Dynamic) return C;
    return R == 1 ? o.size() : o.cols();
  }

  Index rows_() const { return st_.rows(); }
  Index cols_() const { return st_.cols(); }
  Index rstride_() const { return 1; }
  Index cstride_() const { return st_.rows(); }
  S at(Index i, Index j) const { return st_.data()[i + j * st_.rows()]; }
  S& at(Index i, Index j) { return st_.data()[i + j * st_.rows()]; }
  S* data() { return st_.data(); }
  const S* data() const { return st_.data(); }
  void resize(Index r, Index c) { st_.resize(r, c); }
  void resize(Index n) { st_.resize(C == 1 ? n : 1, C == 1 ? 1 : n); }

  static Matrix Zero() {
    Matrix m;
    m.setZero();
    return m;
  }
  static Matrix Zero(Index n) {
    Matrix m(n);
    m.setZero();
    return m;
  }
  static Matrix Zero(Index r, Index c) {
    Matrix m(r, c);
    m.setZero();
    return m;
  }
  static Matrix Identity() {
    Matrix m;
    m.setZero();
    for (Index i = 0; i < std::min(m.rows(), m.cols()); i++) m.at(i, i) = S(1);
    return m;
  }
};

// strided reference into someone else's storage (Block / col / row / head)
template <class S, int R, int C>
struct View : public DenseBase<View<S, R, C>> {
  S* p;
  Index r, c, rs, cs;
  View(S* p_, Index r_, Index c_, Index rs_, Index cs_) : p(p_), r(r_), c(c_), rs(rs_), cs(cs_) {}
  View(const View&) = default;
  Index rows_() const { return r; }
  Index cols_() const { return c; }
  Index rstride_() const { return rs; }
  Index cstride_() const { return cs; }
  S at(Index i, Index j) const { return p[i * rs + j * cs]; }
  S& at(Index i, Index j) { return p[i * rs + j * cs]; }
  View& operator=(const View& o) { return this->assign_from(o); }
  template <class O>
  View& operator=(const DenseBase<O>& o) {
    return this->assign_from(o);
  }
};
template <class M, int R, int C>
using Block = View<typename M::Scalar, R, C>;

// ---- binary operators --------------------------------------------------------------------------
template <class A, class B>
Matrix<typename Traits<A>::Scalar, detail::pick(Traits<A>::Rows, Traits<B>::Rows), detail::pick(Traits<A>::Cols, Traits<B>::Cols)>
operator+(const DenseBase<A>& a, const DenseBase<B>& b) {
  Matrix<typename Traits<A>::Scalar, detail::pick(Traits<A>::Rows, Traits<B>::Rows), detail::pick(Traits<A>::Cols, Traits<B>::Cols)> r(
      a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); j++)
    for (Index i = 0; i < a.rows(); i++) r.at(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return r;
}
template <class A, class B>
Matrix<typename Traits<A>::Scalar, detail::pick(Traits<A>::Rows, Traits<B>::Rows), detail::pick(Traits<A>::Cols, Traits<B>::Cols)>
operator-(const DenseBase<A>& a, const DenseBase<B>& b) {
  Matrix<typename Traits<A>::Scalar, detail::pick(Traits<A>::Rows, Traits<B>::Rows), detail::pick(Traits<A>::Cols, Traits<B>::Cols)> r(
      a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); j++)
    for (Index i = 0; i < a.rows(); i++) r.at(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return r;
}
template <class A, class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>
typename DenseBase<A>::Plain operator*(const DenseBase<A>& a, T s) {
  using S = typename Traits<A>::Scalar;
  return a.unary([s](S v) { return v * S(s); });
}
template <class A, class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>
typename DenseBase<A>::Plain operator*(T s, const DenseBase<A>& a) {
  using S = typename Traits<A>::Scalar;
  return a.unary([s](S v) { return S(s) * v; });
}
template <class A, class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>
typename DenseBase<A>::Plain operator/(const DenseBase<A>& a, T s) {
  using S = typename Traits<A>::Scalar;
  return a.unary([s](S v) { return v / S(s); });
}
// matrix product, coefficient based: (lhs.row(i).cwiseProduct(rhs.col(j))).sum()
template <class A, class B>
Matrix<typename Traits<A>::Scalar, Traits<A>::Rows, Traits<B>::Cols> operator*(const DenseBase<A>& a, const DenseBase<B>& b) {
  using S = typename Traits<A>::Scalar;
  constexpr bool fixed_inner = Traits<A>::Cols != Dynamic || Traits<B>::Rows != Dynamic;
  Matrix<S, Traits<A>::Rows, Traits<B>::Cols> r(a.rows(), b.cols());
  const Index n = a.cols();
  for (Index j = 0; j < b.cols(); j++)
    for (Index i = 0; i < a.rows(); i++)
      r.at(i, j) = n == 0 ? S(0)
                          : detail::redux<fixed_inner>(
                                n, [&](Index k) { return a.coeff(i, k) * b.coeff(k, j); }, [](S x, S y) { return x + y; });
  return r;
}

template <class M>
struct ColwiseProxy {
  const M& m;
  template <class V>
  typename DenseBase<M>::Plain operator+(const DenseBase<V>& v) const {
    typename DenseBase<M>::Plain r(m.rows(), m.cols());
    for (Index j = 0; j < m.cols(); j++)
      for (Index i = 0; i < m.rows(); i++) r.at(i, j) = m.coeff(i, j) + v.lin(i);
    return r;
  }
};
template <class M>
struct RowwiseProxy {
  const M& m;
  using S = typename Traits<M>::Scalar;
  Matrix<S, Traits<M>::Rows, 1> minCoeff() const {
    Matrix<S, Traits<M>::Rows, 1> r(m.rows(), Index(1));
    for (Index i = 0; i < m.rows(); i++) r.at(i, 0) = m.row(i).minCoeff();
    return r;
  }
  Matrix<S, Traits<M>::Rows, 1> maxCoeff() const {
    Matrix<S, Traits<M>::Rows, 1> r(m.rows(), Index(1));
    for (Index i = 0; i < m.rows(); i++) r.at(i, 0) = m.row(i).maxCoeff();
    return r;
  }
};

// ---- ArrayWrapper: coefficient-wise arithmetic on a referenced matrix or on an owned temporary -----
template <class S, int R, int C>
struct ArrayOf {
  using M = Matrix<S, R, C>;
  View<S, R, C> v;  // refers either to the source matrix or to `own`
  M own;
  ArrayOf(const View<S, R, C>& src) : v(src) {}
  explicit ArrayOf(M&& m) : v(nullptr, 0, 0, 0, 0), own(std::move(m)) { rebind(); }
  ArrayOf(const ArrayOf& o) : v(o.v), own(o.own) {
    if (o.v.p == o.own.data()) rebind();
  }
  void rebind() {  // (View::operator= copies coefficients, so the fields are set one by one)
    v.p = own.data();
    v.r = own.rows();
    v.c = own.cols();
    v.rs = own.rstride_();
    v.cs = own.cstride_();
  }
  template <class F>
  ArrayOf map(const F& f) const {
    return ArrayOf(v.unary(f));
  }
  ArrayOf abs() const {
    return map([](S a) { return std::abs(a); });
  }
  ArrayOf max(S s) const {
    return map([s](S a) { return std::max(a, s); });
  }
  ArrayOf min(S s) const {
    return map([s](S a) { return std::min(a, s); });
  }
  ArrayOf operator+(S s) const {
    return map([s](S a) { return a + s; });
  }
  ArrayOf operator-(S s) const {
    return map([s](S a) { return a - s; });
  }
  ArrayOf operator-(const ArrayOf& o) const { return ArrayOf(M(v - o.v)); }
  ArrayOf operator+(const ArrayOf& o) const { return ArrayOf(M(v + o.v)); }
  ArrayOf& operator-=(S s) {
    for (Index k = 0; k < v.size(); k++) v.linRef(k) -= s;
    return *this;
  }
  ArrayOf& operator+=(S s) {
    for (Index k = 0; k < v.size(); k++) v.linRef(k) += s;
    return *this;
  }
  M matrix() const { return M(v); }
};
template <class S, int R, int C>
Matrix<S, R, C>::Matrix(const ArrayOf<S, R, C>& a) : Matrix(a.v) {}
template <class S, int R, int C>
template <int R2, int C2>
Matrix<S, R, C>::Matrix(const ArrayOf<S, R2, C2>& a) : Matrix(a.v) {}
template <class S, int R, int C>
template <int R2, int C2>
Matrix<S, R, C>& Matrix<S, R, C>::operator=(const ArrayOf<S, R2, C2>& a) {
  return *this = a.v;
}

// ---------------------------------------------------------------------------------------------
template <class S>
class Quaternion {
  Matrix<S, 4, 1> c_;  // x, y, z, w

 public:
  using Vector3 = Matrix<S, 3, 1>;
  Quaternion() {}
  Quaternion(S w, S x, S y, S z) : c_(x, y, z, w) {}
  S x() const { return c_[0]; }
  S y() const { return c_[1]; }
  S z() const { return c_[2]; }
  S w() const { return c_[3]; }
  S& x() { return c_[0]; }
  S& y() { return c_[1]; }
  S& z() { return c_[2]; }
  S& w() { return c_[3]; }
  Matrix<S, 4, 1>& coeffs() { return c_; }
  const Matrix<S, 4, 1>& coeffs() const { return c_; }
  static Quaternion Identity() { return Quaternion(S(1), S(0), S(0), S(0)); }
  Quaternion& setIdentity() {
    *this = Identity();
    return *this;
  }
  Vector3 vec() const { return Vector3(x(), y(), z()); }
  S squaredNorm() const { return c_.squaredNorm(); }
  // Quaternion.h quat_product<Arch::None>
  Quaternion operator*(const Quaternion& b) const {
    const Quaternion& a = *this;
    return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                      a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                      a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                      a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
  }
  // QuaternionBase::_transformVector
  template <class V>
  Vector3 operator*(const DenseBase<V>& v) const {
    Vector3 uv = vec().cross(v);
    uv += uv;
    Vector3 wuv = w() * uv;
    Vector3 t = v + wuv;
    return t + vec().cross(uv);
  }
  // QuaternionBase::toRotationMatrix
  Matrix<S, 3, 3> toRotationMatrix() const {
    Matrix<S, 3, 3> res;
    const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
    const S twx = tx * w(), twy = ty * w(), twz = tz * w();
    const S txx = tx * x(), txy = ty * x(), txz = tz * x();
    const S tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    res.at(0, 0) = S(1) - (tyy + tzz);
    res.at(0, 1) = txy - twz;
    res.at(0, 2) = txz + twy;
    res.at(1, 0) = txy + twz;
    res.at(1, 1) = S(1) - (txx + tzz);
    res.at(1, 2) = tyz - twx;
    res.at(2, 0) = txz - twy;
    res.at(2, 1) = tyz + twx;
    res.at(2, 2) = S(1) - (txx + tyy);
    return res;
  }
  // QuaternionBase::inverse: conjugate / squaredNorm
  Quaternion inverse() const {
    const S n2 = squaredNorm();
    if (n2 > S(0)) return Quaternion(w() / n2, -x() / n2, -y() / n2, -z() / n2);
    return Quaternion(S(0), S(0), S(0), S(0));
  }
  template <class T>
  Quaternion<T> cast() const {
    return Quaternion<T>(T(w()), T(x()), T(y()), T(z()));
  }
};

using Vector2d = Matrix<double, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Vector3f = Matrix<float, 3, 1>;
using Matrix3d = Matrix<double, 3, 3>;
using VectorXd = Matrix<double, Dynamic, 1>;
using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using Matrix3Xd = Matrix<double, 3, Dynamic>;
using Quaterniond = Quaternion<double>;
using Quaternionf = Quaternion<float>;

}  // namespace Eigen
