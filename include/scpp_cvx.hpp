// include/scpp_cvx.hpp — a RECORDING shim of the constraint DSL the reference's models and problem builders are written
// against (Epigraph, namespace cvx; submodule lib/Epigraph @ eabeed5, absent here).  Host-only, header-only, C++17.
//
// It covers exactly the subset SURVEY §8(b) lists, i.e. every call made by
//      scpp_models/src/rocketQuat.cpp:70-144, scpp_models/src/rocket2d.cpp:46-84        (addApplicationConstraints)
//      scpp_core/src/SCProblem.cpp:16-134, SCvxProblem.cpp:12-67, MPCProblem.cpp:16-86   (problem builders)
//   cvx::OptimizationProblem { addVariable(name[, rows[, cols]]), getVariable(name, var&), addConstraint, addCostTerm, getVariableValue }
//   cvx::Scalar / VectorX / MatrixX with  + - * (by parameter),  (i,j) (i) col row block topRows rightCols head tail,
//        colwise().norm() colwise().sum() norm() sum() cwiseProduct cols() rows(), the comma initialiser  v << a, b;
//   cvx::par(value) cvx::dynpar(reference)  — a dynpar keeps the ADDRESS, so the value is re-read whenever the problem is evaluated
//        (the reference's pointer semantics, SURVEY §8(b) "Ownership")
//   cvx::equalTo lessThan greaterThan box
// Nothing is solved here: the problem is RECORDED as affine expressions over named variables.  include/scpp_plugin.hpp lowers
// the recorded application constraints of a model to the stage-wise row table / pinned-variable masks the device kernels use
// (scpp_b200/csrc/models.cuh: RowDesc), which is how a model written in the reference's style drops into the engine.
#pragma once
#include <cmath>
#include <cstddef>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace cvx {

// ---- parameters -------------------------------------------------------------------------------------------------------------
// value = scale * (*ptr)  (dynpar)  or  lit  (par / plain number)
struct Param {
    const double *ptr = nullptr;
    double scale = 1., lit = 0.;
    Param() {}
    Param(double v) : lit(v) {}
    static Param dyn(const double *p, double s = 1.) { Param q; q.ptr = p; q.scale = s; return q; }
    bool is_dynamic() const { return ptr != nullptr; }
    double value() const { return ptr ? scale * *ptr : lit; }
    Param operator-() const { Param q = *this; if (ptr) q.scale = -scale; else q.lit = -lit; return q; }
};
inline Param operator*(const Param &a, const Param &b)
{
    if (a.ptr && b.ptr) throw std::runtime_error("cvx shim: product of two dynamic parameters is not in the reference's subset");
    if (a.ptr) return Param::dyn(a.ptr, a.scale * b.lit);
    if (b.ptr) return Param::dyn(b.ptr, b.scale * a.lit);
    return Param(a.lit * b.lit);
}

// ---- affine expressions over the problem's variables --------------------------------------------------------------------
struct Term { Param coef; int var; };
struct Affine {
    std::vector<Term> terms;
    std::vector<Param> consts;      // their sum is the constant part
    Affine() {}
    Affine(double v) { if (v != 0.) consts.push_back(Param(v)); }
    Affine(const Param &p) { consts.push_back(p); }
    static Affine variable(int id) { Affine a; a.terms.push_back({Param(1.), id}); return a; }
    double constant() const { double c = 0; for (auto &p : consts) c += p.value(); return c; }
    double evaluate(const std::vector<double> &x) const { double v = constant(); for (auto &t : terms) v += t.coef.value() * x[t.var]; return v; }
    Affine &operator+=(const Affine &o) { terms.insert(terms.end(), o.terms.begin(), o.terms.end()); consts.insert(consts.end(), o.consts.begin(), o.consts.end()); return *this; }
    Affine operator-() const { Affine r = *this; for (auto &t : r.terms) t.coef = -t.coef; for (auto &p : r.consts) p = -p; return r; }
};
inline Affine operator+(Affine a, const Affine &b) { a += b; return a; }
inline Affine operator-(Affine a, const Affine &b) { a += -b; return a; }
inline Affine operator*(const Param &p, const Affine &a) { Affine r = a; for (auto &t : r.terms) t.coef = p * t.coef; for (auto &c : r.consts) c = p * c; return r; }
inline Affine operator*(const Affine &a, const Param &p) { return p * a; }
inline Affine operator+(const Param &p, const Affine &a) { return Affine(p) + a; }
inline Affine operator+(const Affine &a, const Param &p) { return a + Affine(p); }
inline Affine operator-(const Param &p, const Affine &a) { return Affine(p) - a; }
inline Affine operator-(const Affine &a, const Param &p) { return a - Affine(p); }
inline Affine operator+(const Param &a, const Param &b) { return Affine(a) + Affine(b); }

// ||tail||_2 <= head is recorded from  lessThan(norm expression, affine)
struct Norm2 { std::vector<Affine> tail; };
struct NormRow { std::vector<Norm2> n; };                       // one norm per column: X.block(..).colwise().norm()

// ---- matrices of affine expressions (column-major like Eigen) --------------------------------------------------------------
template <class T>
struct Mat {
    int r = 0, c = 0;
    std::vector<T> v;
    Mat() {}
    explicit Mat(int rows, int cols = 1) : r(rows), c(cols), v((size_t)rows * cols) {}
    Mat(const T &scalar) : r(1), c(1), v(1, scalar) {}
    int rows() const { return r; }
    int cols() const { return c; }
    int size() const { return r * c; }
    T &operator()(int i, int j) { return v[(size_t)j * r + i]; }
    const T &operator()(int i, int j) const { return v[(size_t)j * r + i]; }
    T &operator()(int i) { return v[i]; }
    const T &operator()(int i) const { return v[i]; }
    operator T() const { if (r * c != 1) throw std::runtime_error("cvx shim: matrix used as a scalar"); return v[0]; }
    Mat block(int i0, int j0, int nr, int nc) const { Mat m(nr, nc); for (int j = 0; j < nc; j++) for (int i = 0; i < nr; i++) m(i, j) = (*this)(i0 + i, j0 + j); return m; }
    Mat col(int j) const { return block(0, j, r, 1); }
    Mat row(int i) const { return block(i, 0, 1, c); }
    Mat topRows(int n) const { return block(0, 0, n, c); }
    Mat bottomRows(int n) const { return block(r - n, 0, n, c); }
    Mat leftCols(int n) const { return block(0, 0, r, n); }
    Mat rightCols(int n) const { return block(0, c - n, r, n); }
    Mat head(int n) const { Mat m(n, 1); for (int i = 0; i < n; i++) m(i) = v[i]; return m; }
    Mat tail(int n) const { Mat m(n, 1); for (int i = 0; i < n; i++) m(i) = v[r * c - n + i]; return m; }
    // writable tail (SCProblem.cpp:122: norm2_terms.tail(n) = ...)
    struct TailRef { Mat &m; int n; void operator=(const Mat &o) { for (int i = 0; i < n; i++) m.v[m.v.size() - n + i] = o.v[i]; } };
    TailRef tail_ref(int n) { return TailRef{*this, n}; }
    // comma initialiser  v << a, b, c;   (scalars or blocks, filled in storage order)
    struct Comma {
        Mat &m; int at;
        Comma &operator,(const T &x) { m.v.at(at++) = x; return *this; }
        Comma &operator,(const Mat &x) { for (auto &e : x.v) m.v.at(at++) = e; return *this; }
    };
    Comma operator<<(const T &x) { v.at(0) = x; return Comma{*this, 1}; }
    Comma operator<<(const Mat &x) { int a = 0; for (auto &e : x.v) v.at(a++) = e; return Comma{*this, a}; }
    // reductions in the reference's member syntax (instantiated only for matrices of affine expressions)
    struct ColwiseProxy {
        const Mat &m;
        NormRow norm() const { NormRow r; for (int j = 0; j < m.c; j++) { Norm2 q; for (int i = 0; i < m.r; i++) q.tail.push_back(m(i, j)); r.n.push_back(q); } return r; }
        Mat sum() const { Mat s(1, m.c); for (int j = 0; j < m.c; j++) { T a; for (int i = 0; i < m.r; i++) a += m(i, j); s(0, j) = a; } return s; }
    };
    ColwiseProxy colwise() const { return ColwiseProxy{*this}; }
    Norm2 norm() const { Norm2 q; q.tail = v; return q; }
    T sum() const { T a; for (auto &e : v) a += e; return a; }
    template <class B> auto cwiseProduct(const Mat<B> &b) const;      // parameter matrix .cwiseProduct(variable matrix), rocketQuat.cpp:119
};
using MatrixX = Mat<Affine>;
using VectorX = Mat<Affine>;
using Scalar = Mat<Affine>;
using ParamMat = Mat<Param>;

// elementwise algebra (with broadcasting of 1x1 operands)
template <class A, class B, class F>
auto zip(const Mat<A> &a, const Mat<B> &b, F f) -> Mat<decltype(f(a.v[0], b.v[0]))>
{
    using R = decltype(f(a.v[0], b.v[0]));
    const bool sa = a.size() == 1, sb = b.size() == 1;
    if (!sa && !sb && (a.r != b.r || a.c != b.c)) throw std::runtime_error("cvx shim: shape mismatch");
    Mat<R> m(sa ? b.r : a.r, sa ? b.c : a.c);
    for (int i = 0; i < m.size(); i++) m.v[i] = f(a.v[sa ? 0 : i], b.v[sb ? 0 : i]);
    return m;
}
inline MatrixX operator+(const MatrixX &a, const MatrixX &b) { return zip(a, b, [](const Affine &x, const Affine &y) { return x + y; }); }
inline MatrixX operator-(const MatrixX &a, const MatrixX &b) { return zip(a, b, [](const Affine &x, const Affine &y) { return x - y; }); }
inline MatrixX operator+(const ParamMat &a, const MatrixX &b) { return zip(a, b, [](const Param &x, const Affine &y) { return Affine(x) + y; }); }
inline MatrixX operator+(const MatrixX &a, const ParamMat &b) { return b + a; }
inline MatrixX operator-(const ParamMat &a, const MatrixX &b) { return zip(a, b, [](const Param &x, const Affine &y) { return Affine(x) - y; }); }
inline MatrixX operator-(const MatrixX &a, const ParamMat &b) { return zip(a, b, [](const Affine &x, const Param &y) { return x - Affine(y); }); }
inline MatrixX operator+(const ParamMat &a, const ParamMat &b) { return zip(a, b, [](const Param &x, const Param &y) { return Affine(x) + Affine(y); }); }
inline MatrixX operator-(const MatrixX &a) { MatrixX m = a; for (auto &e : m.v) e = -e; return m; }
inline ParamMat operator-(const ParamMat &a) { ParamMat m = a; for (auto &e : m.v) e = -e; return m; }
inline MatrixX &operator+=(MatrixX &a, const MatrixX &b) { a = a + b; return a; }
// parameter (matrix) * variable (matrix): scalar scaling or the matrix product  A_k * x_k  of the dynamics rows (SCProblem.cpp:44-55)
inline MatrixX operator*(const ParamMat &a, const MatrixX &b)
{
    if (a.size() == 1 || b.size() == 1) return zip(a, b, [](const Param &x, const Affine &y) { return x * y; });
    if (a.c != b.r) throw std::runtime_error("cvx shim: matrix product shape mismatch");
    MatrixX m(a.r, b.c);
    for (int i = 0; i < a.r; i++) for (int j = 0; j < b.c; j++) { Affine s; for (int q = 0; q < a.c; q++) s += a(i, q) * b(q, j); m(i, j) = s; }
    return m;
}
inline MatrixX operator*(const MatrixX &b, const ParamMat &a) { if (a.size() != 1) throw std::runtime_error("cvx shim: variable * parameter matrix"); return a * b; }
inline MatrixX operator*(const Param &p, const MatrixX &b) { return ParamMat(p) * b; }
inline Affine operator*(const ParamMat &p, const Affine &a) { return (Param)p * a; }      // dynpar(weight) * delta.sum()  (SCProblem.cpp:134)
inline Affine operator*(const Affine &a, const ParamMat &p) { return (Param)p * a; }

template <class T> template <class B> auto Mat<T>::cwiseProduct(const Mat<B> &b) const { return zip(*this, b, [](const T &x, const B &y) { return x * y; }); }

// ---- par / dynpar -------------------------------------------------------------------------------------------------------
inline ParamMat par(double v) { return ParamMat(Param(v)); }
inline ParamMat dynpar(const double &v) { return ParamMat(Param::dyn(&v)); }
// contiguous storage (std::vector, std::array, C arrays, the small-matrix types of a model): column vector, or rows x cols column-major
inline ParamMat dynpar(const double *data, int rows, int cols = 1) { ParamMat m(rows, cols); for (int i = 0; i < rows * cols; i++) m.v[i] = Param::dyn(data + i); return m; }
template <class C> auto dynpar(const C &c) -> decltype(c.data(), c.size(), ParamMat()) { return dynpar(c.data(), (int)c.size()); }
template <size_t N> ParamMat dynpar(const double (&a)[N]) { return dynpar(a, (int)N); }

// ---- constraints --------------------------------------------------------------------------------------------------------
struct Constraint {
    enum Kind { EQ, LE, SOC } kind;       // EQ: lhs == rhs ;  LE: lhs <= rhs ;  SOC: ||tail|| <= rhs
    Affine lhs, rhs;
    std::vector<Affine> tail;
};
using ConstraintList = std::vector<Constraint>;
inline MatrixX lift(const ParamMat &p) { MatrixX m(p.r, p.c); for (int i = 0; i < p.size(); i++) m.v[i] = Affine(p.v[i]); return m; }
inline MatrixX lift(double v) { return MatrixX(Affine(v)); }
inline const MatrixX &lift(const MatrixX &m) { return m; }
inline ConstraintList relate(Constraint::Kind k, const MatrixX &a, const MatrixX &b)
{
    ConstraintList l;
    auto z = zip(a, b, [&](const Affine &x, const Affine &y) { Constraint c; c.kind = k; c.lhs = x; c.rhs = y; return c; });
    return z.v;
}
template <class A, class B> ConstraintList equalTo(const A &a, const B &b) { return relate(Constraint::EQ, lift(a), lift(b)); }
template <class A, class B> ConstraintList greaterThan(const A &a, const B &b) { return relate(Constraint::LE, lift(b), lift(a)); }
template <class A, class B> ConstraintList lessThan(const A &a, const B &b) { return relate(Constraint::LE, lift(a), lift(b)); }
template <class B> ConstraintList lessThan(const Norm2 &n, const B &b)
{
    const MatrixX h = lift(b);
    if (h.size() != 1) throw std::runtime_error("cvx shim: norm <= non-scalar");
    Constraint c; c.kind = Constraint::SOC; c.tail = n.tail; c.rhs = h.v[0];
    return {c};
}
template <class B> ConstraintList lessThan(const NormRow &n, const B &b)
{
    const MatrixX h = lift(b);
    ConstraintList l;
    for (size_t j = 0; j < n.n.size(); j++) {
        Constraint c; c.kind = Constraint::SOC; c.tail = n.n[j].tail; c.rhs = h.size() == 1 ? h.v[0] : h.v.at(j);
        l.push_back(c);
    }
    return l;
}
template <class L, class X, class U> ConstraintList box(const L &lo, const X &x, const U &hi)
{
    ConstraintList l = relate(Constraint::LE, lift(lo), lift(x)), u = relate(Constraint::LE, lift(x), lift(hi));
    l.insert(l.end(), u.begin(), u.end());
    return l;
}

// ---- the problem ------------------------------------------------------------------------------------------------------------
class OptimizationProblem {
public:
    struct Var { std::string name; int offset, rows, cols; };
    MatrixX addVariable(const std::string &name, int rows = 1, int cols = 1)
    {
        if (index_.count(name)) throw std::runtime_error("cvx shim: variable '" + name + "' exists");
        index_[name] = (int)vars_.size();
        vars_.push_back({name, n_, rows, cols});
        MatrixX m(rows, cols);
        for (int i = 0; i < rows * cols; i++) m.v[i] = Affine::variable(n_ + i);
        n_ += rows * cols;
        return m;
    }
    void getVariable(const std::string &name, MatrixX &out) const
    {
        const Var &v = var(name);
        out = MatrixX(v.rows, v.cols);
        for (int i = 0; i < v.rows * v.cols; i++) out.v[i] = Affine::variable(v.offset + i);
    }
    void addConstraint(const ConstraintList &l) { constraints.insert(constraints.end(), l.begin(), l.end()); }
    void addCostTerm(const Affine &a) { cost += a; }
    void addCostTerm(const MatrixX &a) { cost += (Affine)a; }
    const Var &var(const std::string &name) const
    {
        auto it = index_.find(name);
        if (it == index_.end()) throw std::runtime_error("cvx shim: no variable '" + name + "'");
        return vars_[it->second];
    }
    int numVariables() const { return n_; }
    const std::vector<Var> &variables() const { return vars_; }
    // values of a variable in a solution vector x (what socp->getVariableValue reads after a solve, SCAlgorithm.cpp:195-200)
    std::vector<double> getVariableValue(const std::string &name, const std::vector<double> &x) const
    {
        const Var &v = var(name);
        return std::vector<double>(x.begin() + v.offset, x.begin() + v.offset + v.rows * v.cols);
    }
    // residuals of every recorded constraint at x: EQ |lhs - rhs|, LE max(lhs - rhs, 0), SOC max(||tail|| - rhs, 0)
    double violation(const std::vector<double> &x) const
    {
        double worst = 0;
        for (auto &c : constraints) {
            double v;
            if (c.kind == Constraint::EQ) v = std::fabs(c.lhs.evaluate(x) - c.rhs.evaluate(x));
            else if (c.kind == Constraint::LE) v = c.lhs.evaluate(x) - c.rhs.evaluate(x);
            else { double n = 0; for (auto &t : c.tail) { const double e = t.evaluate(x); n += e * e; } v = std::sqrt(n) - c.rhs.evaluate(x); }
            if (v > worst) worst = v;
        }
        return worst;
    }
    ConstraintList constraints;
    Affine cost;

private:
    std::vector<Var> vars_;
    std::map<std::string, int> index_;
    int n_ = 0;
};

} // namespace cvx
