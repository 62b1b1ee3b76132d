// TMB.hpp -- a minimal stand-in for the TMB / Eigen / R surface that the UNMODIFIED reference
// templates use, so that /root/reference/src/smoothSDE.cpp and src/nllk/*.hpp compile here
// with plain g++ (TMB, RcppEigen and R are not installed in this image).
//
// TEST INFRASTRUCTURE ONLY.  Nothing under smoothsde_b200/ includes, links or loads this; only
// tests/, __graft_entry__.smoke() and bench.py's CPU legs use the library built from it
// (oracle/_ref/libsmoothsde_ref.so, recipe: oracle/Makefile, driver: oracle/ref_driver.cpp).
//
// What is the reference's and what is ours: the model code (the objective: data handling, linear
// predictor, Kalman recursions, transition densities, penalties) is the reference's own source,
// compiled where it lies.  This header supplies what the reference gets from third parties:
//   * containers with Eigen's semantics for the members the templates call: `vector<Type>`
//     (Eigen::Array, coefficient-wise `*`), `matrix<Type>` (Eigen::Matrix, `*` = product),
//     `Eigen::SparseMatrix<Type>`, `tmbutils::array<Type>`; all eager, no expression templates;
//   * the TMB macros DATA_* / PARAMETER* / REPORT reading from a name -> buffer table;
//   * `dnorm`, `dt`, `besselI`, `density::GMRF(Q).Quadform`, `atomic::matinvpd`, `atomic::logdet`,
//     `asSparseMatrix`, `diff`, `R_IsNA`, `asDouble`, `error` as documented by TMB / R;
//   * an AD scalar: `ad::Var<Base>`, a reverse-mode tape (what CppAD / TMBad do for TMB: the
//     gradient is one reverse sweep over the recorded operations).  `Base = double` gives the
//     gradient, `Base = ad::Dual` (one forward tangent) gives exact Hessian-vector products by
//     reverse-over-forward, the stand-in for TMB's AD-of-AD Hessian.
#ifndef SSDE_TMB_SHIM_HPP
#define SSDE_TMB_SHIM_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

// ------------------------------------------------------------------------------------------
// AD scalars
// ------------------------------------------------------------------------------------------
namespace ad {
using std::exp;
using std::log;
using std::sqrt;

// forward-mode pair (value, derivative along one direction)
struct Dual {
    double v, d;
    Dual() : v(0), d(0) {}
    Dual(double v_) : v(v_), d(0) {}
    Dual(double v_, double d_) : v(v_), d(d_) {}
};
inline Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
inline Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
inline Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
inline Dual operator/(Dual a, Dual b) { double q = a.v / b.v; return Dual(q, (a.d - q * b.d) / b.v); }
inline Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
inline Dual& operator+=(Dual& a, Dual b) { a = a + b; return a; }
inline Dual exp(Dual a) { double e = std::exp(a.v); return Dual(e, e * a.d); }
inline Dual log(Dual a) { return Dual(std::log(a.v), a.d / a.v); }
inline Dual sqrt(Dual a) { double s = std::sqrt(a.v); return Dual(s, 0.5 * a.d / s); }
inline bool is_zero(double x) { return x == 0.0; }
inline bool is_zero(Dual x) { return x.v == 0.0 && x.d == 0.0; }
inline double value_of(double x) { return x; }
inline double value_of(Dual x) { return x.v; }

inline double digamma(double x) {
    double r = 0;
    while (x < 6) { r -= 1 / x; x += 1; }
    double f = 1 / (x * x);
    return r + std::log(x) - 0.5 / x - f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f / 132))));
}
inline double trigamma(double x) {
    double r = 0;
    while (x < 6) { r += 1 / (x * x); x += 1; }
    double f = 1 / (x * x);
    return r + 1 / x + f / 2 + f / x * (1.0 / 6 - f * (1.0 / 30 - f * (1.0 / 42 - f / 30)));
}
inline double lgamma_(double x) { return std::lgamma(x); }
inline Dual lgamma_(Dual x) { return Dual(std::lgamma(x.v), digamma(x.v) * x.d); }
inline double digamma_(double x) { return digamma(x); }
inline Dual digamma_(Dual x) { return Dual(digamma(x.v), trigamma(x.v) * x.d); }
inline double bessel_i(double x, double nu) { return std::cyl_bessel_i(nu, x); }
// d/dx I_nu(x) = (I_{nu-1} + I_{nu+1}) / 2; d/dnu by a central difference (out-of-scope model)
inline double bessel_i_dx(double x, double nu) { return 0.5 * (std::cyl_bessel_i(std::fabs(nu - 1), x) + std::cyl_bessel_i(nu + 1, x)); }
inline double bessel_i_dnu(double x, double nu) { double h = 1e-6 * (1 + std::fabs(nu)); return (std::cyl_bessel_i(nu + h, x) - std::cyl_bessel_i(nu - h, x)) / (2 * h); }

// reverse-mode tape.  Node 0 is a sink for "no operand"; constants carry index 0.
template <class B>
struct Tape {
    struct Node { uint32_t a, b; B da, db; };
    std::vector<Node> nodes;
    bool recording = false;
    Tape() { nodes.reserve(1 << 20); }
    void reset() { nodes.clear(); nodes.push_back(Node{0, 0, B(0.0), B(0.0)}); }
    uint32_t push(uint32_t a, const B& da, uint32_t b, const B& db) {
        if (nodes.size() >= 0xfffffff0u) throw std::runtime_error("AD tape full");
        nodes.push_back(Node{a, b, da, db});
        return (uint32_t)(nodes.size() - 1);
    }
    static Tape& get() { static thread_local Tape t; return t; }
};

template <class B>
struct Var {
    B v;
    uint32_t i;
    Var() : v(0.0), i(0) {}
    Var(const B& v_) : v(v_), i(0) {}
    template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value && !std::is_same<S, B>::value>::type>
    Var(S s) : v((double)s), i(0) {}
    Var(const B& v_, uint32_t i_) : v(v_), i(i_) {}
    static Var independent(const B& v_) { return Var(v_, Tape<B>::get().push(0, B(0.0), 0, B(0.0))); }
};

template <class B> inline Var<B> mk(const B& v, uint32_t a, const B& da, uint32_t b, const B& db) {
    if (a == 0 && b == 0) return Var<B>(v);
    return Var<B>(v, Tape<B>::get().push(a, da, b, db));
}
template <class B> inline Var<B> operator+(const Var<B>& x, const Var<B>& y) { return mk<B>(x.v + y.v, x.i, B(1.0), y.i, B(1.0)); }
template <class B> inline Var<B> operator-(const Var<B>& x, const Var<B>& y) { return mk<B>(x.v - y.v, x.i, B(1.0), y.i, B(-1.0)); }
template <class B> inline Var<B> operator*(const Var<B>& x, const Var<B>& y) { return mk<B>(x.v * y.v, x.i, y.v, y.i, x.v); }
template <class B> inline Var<B> operator/(const Var<B>& x, const Var<B>& y) {
    B q = x.v / y.v, r = B(1.0) / y.v;
    return mk<B>(q, x.i, r, y.i, -(q * r));
}
template <class B> inline Var<B> operator-(const Var<B>& x) { return mk<B>(-x.v, x.i, B(-1.0), 0, B(0.0)); }
template <class B> inline Var<B> operator+(const Var<B>& x) { return x; }
template <class B> inline Var<B> exp(const Var<B>& x) { B e = exp(x.v); return mk<B>(e, x.i, e, 0, B(0.0)); }
template <class B> inline Var<B> log(const Var<B>& x) { return mk<B>(log(x.v), x.i, B(1.0) / x.v, 0, B(0.0)); }
template <class B> inline Var<B> sqrt(const Var<B>& x) { B s = sqrt(x.v); return mk<B>(s, x.i, B(0.5) / s, 0, B(0.0)); }
template <class B> inline Var<B> lgamma(const Var<B>& x) { return mk<B>(lgamma_(x.v), x.i, digamma_(x.v), 0, B(0.0)); }
template <class B> inline double value_of(const Var<B>& x) { return value_of(x.v); }

// mixed with plain numbers
#define SSDE_AD_MIXED(OP)                                                                                   \
    template <class B, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>        \
    inline Var<B> operator OP(const Var<B>& x, S s) { return x OP Var<B>(B((double)s)); }                   \
    template <class B, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>        \
    inline Var<B> operator OP(S s, const Var<B>& x) { return Var<B>(B((double)s)) OP x; }
SSDE_AD_MIXED(+) SSDE_AD_MIXED(-) SSDE_AD_MIXED(*) SSDE_AD_MIXED(/)
#undef SSDE_AD_MIXED
template <class B> inline Var<B>& operator+=(Var<B>& x, const Var<B>& y) { x = x + y; return x; }
template <class B> inline Var<B>& operator-=(Var<B>& x, const Var<B>& y) { x = x - y; return x; }
template <class B> inline Var<B>& operator*=(Var<B>& x, const Var<B>& y) { x = x * y; return x; }
template <class B> inline Var<B>& operator/=(Var<B>& x, const Var<B>& y) { x = x / y; return x; }

#define SSDE_AD_CMP(OP)                                                                                     \
    template <class B> inline bool operator OP(const Var<B>& x, const Var<B>& y) { return value_of(x) OP value_of(y); } \
    template <class B, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>        \
    inline bool operator OP(const Var<B>& x, S s) { return value_of(x) OP (double)s; }                      \
    template <class B, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>        \
    inline bool operator OP(S s, const Var<B>& x) { return (double)s OP value_of(x); }
SSDE_AD_CMP(<) SSDE_AD_CMP(<=) SSDE_AD_CMP(>) SSDE_AD_CMP(>=) SSDE_AD_CMP(==) SSDE_AD_CMP(!=)
#undef SSDE_AD_CMP

template <class B> inline Var<B> besselI(const Var<B>& x, const Var<B>& nu) {
    double xv = value_of(x), nv = value_of(nu);
    return mk<B>(B(bessel_i(xv, nv)), x.i, B(bessel_i_dx(xv, nv)), nu.i, B(bessel_i_dnu(xv, nv)));
}

// reverse sweep: adjoint of every node given d(result)/d(result) = 1
template <class B>
inline void reverse(const Var<B>& y, std::vector<B>& adj) {
    auto& t = Tape<B>::get();
    adj.assign(t.nodes.size(), B(0.0));
    if (y.i == 0) return;
    adj[y.i] = B(1.0);
    for (size_t k = y.i; k > 0; --k) {
        const auto& n = t.nodes[k];
        const B w = adj[k];
        // "absolute zero" multiply, as CppAD's azmul in its reverse sweeps: an operation the result
        // does not depend on (w == 0) contributes nothing even if its partial is inf / NaN.  The
        // reference needs it: the prediction at a track's last row uses the cross-track dt
        // (nllk_ctcrw.hpp:126-129,206-208; negative when the next track's clock restarts, so
        // exp(-beta dt) overflows) and is then discarded at the ID change (:196-200).
        if (is_zero(w)) continue;
        adj[n.a] += n.da * w;
        adj[n.b] += n.db * w;
    }
}
}  // namespace ad

using ad::exp;
using ad::log;
using ad::sqrt;
using ad::lgamma;

// ------------------------------------------------------------------------------------------
// R API bits
// ------------------------------------------------------------------------------------------
#ifndef M_PI
#define M_PI 3.141592653589793238462643383280
#endif
// R's NA_real_ is the NaN whose low word is 1954 (R: src/main/arithmetic.c); a plain NaN is not NA.
inline int R_IsNA(double x) {
    if (!std::isnan(x)) return 0;
    uint64_t u;
    std::memcpy(&u, &x, 8);
    return (uint32_t)(u & 0xffffffffu) == 1954u;
}
inline double asDouble(double x) { return x; }
inline double asDouble(const ad::Dual& x) { return x.v; }
template <class B> inline double asDouble(const ad::Var<B>& x) { return ad::value_of(x); }
[[noreturn]] inline void error(const char* msg) { throw std::runtime_error(msg); }

template <class T> struct isDouble { enum { value = 0 }; };
template <> struct isDouble<double> { enum { value = 1 }; };

// ------------------------------------------------------------------------------------------
// containers
// ------------------------------------------------------------------------------------------
template <class T> struct matrix;
template <class T> struct vector;
namespace Eigen { template <class T> struct SparseMatrix; }

namespace shim_detail {
template <class S, class T> struct is_scalar_for {
    static const bool value = std::is_arithmetic<S>::value || std::is_same<S, T>::value;
};
template <class T> struct RowRef;
template <class T> struct ColRef;
}  // namespace shim_detail

// Eigen::Array<Type, Dynamic, 1>
template <class T>
struct vector {
    std::vector<T> x;
    vector() {}
    vector(int n) : x((size_t)n) {}
    vector(long n) : x((size_t)n) {}
    vector(size_t n) : x(n) {}
    vector(const std::vector<T>& v) : x(v) {}
    vector(const matrix<T>& m);                 // a row or a column
    vector(const shim_detail::RowRef<T>& r);
    vector(const shim_detail::ColRef<T>& c);
    int size() const { return (int)x.size(); }
    T& operator()(int i) { return x[(size_t)i]; }
    const T& operator()(int i) const { return x[(size_t)i]; }
    T& operator[](int i) { return x[(size_t)i]; }
    const T& operator[](int i) const { return x[(size_t)i]; }
    void setZero() { for (auto& e : x) e = T(0.0); }
    void resize(int n) { x.resize((size_t)n); }
    vector segment(int start, int len) const {
        vector r(len);
        for (int i = 0; i < len; i++) r.x[i] = x[(size_t)(start + i)];
        return r;
    }
    T sum() const { T s(0.0); for (const auto& e : x) s = s + e; return s; }
    vector array() const { return *this; }
    matrix<T> matrix_() const;              // Eigen's .matrix(): a column
};

#define SSDE_VEC_BIN(OP)                                                                                    \
    template <class T> inline vector<T> operator OP(const vector<T>& a, const vector<T>& b) {               \
        if (a.size() != b.size()) throw std::runtime_error("vector size mismatch");                        \
        vector<T> r(a.size());                                                                              \
        for (int i = 0; i < a.size(); i++) r.x[i] = a.x[i] OP b.x[i];                                       \
        return r;                                                                                           \
    }                                                                                                       \
    template <class T, class S, class = typename std::enable_if<shim_detail::is_scalar_for<S, T>::value>::type> \
    inline vector<T> operator OP(const vector<T>& a, const S& s) {                                          \
        vector<T> r(a.size());                                                                              \
        for (int i = 0; i < a.size(); i++) r.x[i] = a.x[i] OP s;                                            \
        return r;                                                                                           \
    }                                                                                                       \
    template <class T, class S, class = typename std::enable_if<shim_detail::is_scalar_for<S, T>::value>::type> \
    inline vector<T> operator OP(const S& s, const vector<T>& a) {                                          \
        vector<T> r(a.size());                                                                              \
        for (int i = 0; i < a.size(); i++) r.x[i] = s OP a.x[i];                                            \
        return r;                                                                                           \
    }
SSDE_VEC_BIN(+) SSDE_VEC_BIN(-) SSDE_VEC_BIN(*) SSDE_VEC_BIN(/)
#undef SSDE_VEC_BIN
template <class T> inline vector<T> operator-(const vector<T>& a) {
    vector<T> r(a.size());
    for (int i = 0; i < a.size(); i++) r.x[i] = -a.x[i];
    return r;
}
#define SSDE_VEC_FUN(F)                                                                                     \
    template <class T> inline vector<T> F(const vector<T>& a) {                                             \
        vector<T> r(a.size());                                                                              \
        for (int i = 0; i < a.size(); i++) r.x[i] = F(a.x[i]);                                              \
        return r;                                                                                           \
    }
SSDE_VEC_FUN(exp) SSDE_VEC_FUN(log) SSDE_VEC_FUN(sqrt)
#undef SSDE_VEC_FUN
template <class T> inline vector<T> diff(const vector<T>& a) {
    vector<T> r(a.size() > 0 ? a.size() - 1 : 0);
    for (int i = 0; i + 1 < a.size(); i++) r.x[i] = a.x[i + 1] - a.x[i];
    return r;
}

namespace shim_detail {
template <class T>
struct RowRef {
    matrix<T>* m;
    int i;
    RowRef& operator=(const vector<T>& v);
    RowRef& operator=(const RowRef& o) { return *this = vector<T>(o); }
    matrix<T> transpose() const;            // column
    vector<T> array() const { return vector<T>(*this); }
};
template <class T>
struct ColRef {
    matrix<T>* m;
    int j;
    ColRef& operator=(const vector<T>& v);
    ColRef& operator=(const ColRef& o) { return *this = vector<T>(o); }
    vector<T> array() const { return vector<T>(*this); }
    matrix<T> matrix_() const;
};
// what matrix::array() returns: a 2-D coefficient-wise view; here only ever assigned back
template <class T> struct Array2 { matrix<T> m; };
}  // namespace shim_detail

// Eigen::Matrix<Type, Dynamic, Dynamic>, column-major
template <class T>
struct matrix {
    std::vector<T> x;
    int nr, nc;
    matrix() : nr(0), nc(0) {}
    matrix(int r, int c) : x((size_t)r * (size_t)c), nr(r), nc(c) {}
    matrix(const shim_detail::Array2<T>& a) : x(a.m.x), nr(a.m.nr), nc(a.m.nc) {}
    matrix(const Eigen::SparseMatrix<T>& s);
    int rows() const { return nr; }
    int cols() const { return nc; }
    int size() const { return nr * nc; }
    T& operator()(int i, int j) { return x[(size_t)j * nr + i]; }
    const T& operator()(int i, int j) const { return x[(size_t)j * nr + i]; }
    void setZero() { for (auto& e : x) e = T(0.0); }
    void setIdentity() { setZero(); for (int i = 0; i < nr && i < nc; i++) (*this)(i, i) = T(1.0); }
    shim_detail::RowRef<T> row(int i) { return shim_detail::RowRef<T>{this, i}; }
    shim_detail::ColRef<T> col(int j) { return shim_detail::ColRef<T>{this, j}; }
    shim_detail::RowRef<T> row(int i) const { return shim_detail::RowRef<T>{const_cast<matrix*>(this), i}; }
    shim_detail::ColRef<T> col(int j) const { return shim_detail::ColRef<T>{const_cast<matrix*>(this), j}; }
    matrix block(int r0, int c0, int r, int c) const {
        matrix b(r, c);
        for (int j = 0; j < c; j++) for (int i = 0; i < r; i++) b(i, j) = (*this)(r0 + i, c0 + j);
        return b;
    }
    shim_detail::Array2<T> array() const { return shim_detail::Array2<T>{*this}; }
    matrix transpose() const {
        matrix t(nc, nr);
        for (int j = 0; j < nc; j++) for (int i = 0; i < nr; i++) t(j, i) = (*this)(i, j);
        return t;
    }
    // Eigen's dynamic-size inverse(): LU with partial pivoting
    matrix inverse() const {
        if (nr != nc) throw std::runtime_error("inverse of a non-square matrix");
        int n = nr;
        matrix a(*this), inv(n, n);
        inv.setIdentity();
        for (int c = 0; c < n; c++) {
            int p = c;
            double best = std::fabs(asDouble(a(c, c)));
            for (int r = c + 1; r < n; r++) {
                double v = std::fabs(asDouble(a(r, c)));
                if (v > best) { best = v; p = r; }
            }
            if (p != c) for (int j = 0; j < n; j++) { std::swap(a(c, j), a(p, j)); std::swap(inv(c, j), inv(p, j)); }
            T piv = T(1.0) / a(c, c);
            for (int j = 0; j < n; j++) { a(c, j) = a(c, j) * piv; inv(c, j) = inv(c, j) * piv; }
            for (int r = 0; r < n; r++) {
                if (r == c) continue;
                T f = a(r, c);      // no zero-skipping: a zero-VALUED entry can still be an AD variable
                for (int j = 0; j < n; j++) { a(r, j) = a(r, j) - f * a(c, j); inv(r, j) = inv(r, j) - f * inv(c, j); }
            }
        }
        return inv;
    }
};

template <class T> vector<T>::vector(const ::matrix<T>& m) : x(m.x) {
    if (m.nr != 1 && m.nc != 1 && m.size() != 0) throw std::runtime_error("matrix -> vector needs one row or one column");
}
template <class T> vector<T>::vector(const shim_detail::RowRef<T>& r) : x((size_t)r.m->nc) {
    for (int j = 0; j < r.m->nc; j++) x[j] = (*r.m)(r.i, j);
}
template <class T> vector<T>::vector(const shim_detail::ColRef<T>& c) : x((size_t)c.m->nr) {
    for (int i = 0; i < c.m->nr; i++) x[i] = (*c.m)(i, c.j);
}
template <class T> ::matrix<T> vector<T>::matrix_() const {
    ::matrix<T> m(size(), 1);
    m.x = x;
    return m;
}
namespace shim_detail {
template <class T> RowRef<T>& RowRef<T>::operator=(const vector<T>& v) {
    if (v.size() != m->nc) throw std::runtime_error("row assignment size mismatch");
    for (int j = 0; j < m->nc; j++) (*m)(i, j) = v.x[j];
    return *this;
}
template <class T> ::matrix<T> RowRef<T>::transpose() const {
    ::matrix<T> c(m->nc, 1);
    for (int j = 0; j < m->nc; j++) c(j, 0) = (*m)(i, j);
    return c;
}
template <class T> ColRef<T>& ColRef<T>::operator=(const vector<T>& v) {
    if (v.size() != m->nr) throw std::runtime_error("column assignment size mismatch");
    for (int k = 0; k < m->nr; k++) (*m)(k, j) = v.x[k];
    return *this;
}
template <class T> ::matrix<T> ColRef<T>::matrix_() const { return vector<T>(*this).matrix_(); }
// row * scalar  (nllk_bm_ssm.hpp: `mu.row(i) * dtimes(i)`)
template <class T, class S, class = typename std::enable_if<is_scalar_for<S, T>::value>::type>
inline ::matrix<T> operator*(const RowRef<T>& r, const S& s) {
    ::matrix<T> o(1, r.m->nc);
    for (int j = 0; j < r.m->nc; j++) o(0, j) = (*r.m)(r.i, j) * s;
    return o;
}
}  // namespace shim_detail

template <class T> inline matrix<T> operator+(const matrix<T>& a, const matrix<T>& b) {
    if (a.nr != b.nr || a.nc != b.nc) throw std::runtime_error("matrix + size mismatch");
    matrix<T> r(a.nr, a.nc);
    for (size_t k = 0; k < a.x.size(); k++) r.x[k] = a.x[k] + b.x[k];
    return r;
}
template <class T> inline matrix<T> operator-(const matrix<T>& a, const matrix<T>& b) {
    if (a.nr != b.nr || a.nc != b.nc) throw std::runtime_error("matrix - size mismatch");
    matrix<T> r(a.nr, a.nc);
    for (size_t k = 0; k < a.x.size(); k++) r.x[k] = a.x[k] - b.x[k];
    return r;
}
template <class T> inline matrix<T> operator*(const matrix<T>& a, const matrix<T>& b) {
    if (a.nc != b.nr) throw std::runtime_error("matrix product size mismatch");
    matrix<T> r(a.nr, b.nc);
    for (int j = 0; j < b.nc; j++)
        for (int i = 0; i < a.nr; i++) {
            T s = a(i, 0) * b(0, j);
            for (int k = 1; k < a.nc; k++) s = s + a(i, k) * b(k, j);
            r(i, j) = s;
        }
    return r;
}
template <class T, class S, class = typename std::enable_if<shim_detail::is_scalar_for<S, T>::value>::type>
inline matrix<T> operator*(const matrix<T>& a, const S& s) {
    matrix<T> r(a.nr, a.nc);
    for (size_t k = 0; k < a.x.size(); k++) r.x[k] = a.x[k] * s;
    return r;
}
template <class T, class S, class = typename std::enable_if<shim_detail::is_scalar_for<S, T>::value>::type>
inline matrix<T> operator*(const S& s, const matrix<T>& a) {
    matrix<T> r(a.nr, a.nc);
    for (size_t k = 0; k < a.x.size(); k++) r.x[k] = s * a.x[k];
    return r;
}
// TMB: matrix<Type> * vector<Type> -> vector<Type>
template <class T> inline vector<T> operator*(const matrix<T>& a, const vector<T>& v) {
    if (a.nc != v.size()) throw std::runtime_error("matrix * vector size mismatch");
    vector<T> r(a.nr);
    for (int i = 0; i < a.nr; i++) {
        T s(0.0);
        for (int k = 0; k < a.nc; k++) s = (k == 0) ? a(i, 0) * v.x[0] : s + a(i, k) * v.x[k];
        r.x[i] = s;
    }
    return r;
}

namespace tmbutils {
// tmbutils::array<Type>: column-major with a dim attribute; .col(i) = slice of the LAST dimension
template <class T>
struct array {
    std::vector<T> x;
    std::vector<int> dim;
    int size() const { return (int)x.size(); }
    array col(int i) const {
        array r;
        size_t inner = 1;
        for (size_t k = 0; k + 1 < dim.size(); k++) { inner *= (size_t)dim[k]; r.dim.push_back(dim[k]); }
        r.x.assign(x.begin() + (size_t)i * inner, x.begin() + (size_t)(i + 1) * inner);
        return r;
    }
    ::matrix<T> matrix() const {
        if (dim.size() != 2) throw std::runtime_error("array::matrix() needs a 2-D array");
        ::matrix<T> m(dim[0], dim[1]);
        m.x = x;
        return m;
    }
};
}  // namespace tmbutils
using tmbutils::array;

namespace Eigen {
// compressed-column storage; only what the templates touch
template <class T>
struct SparseMatrix {
    int nr = 0, nc = 0;
    std::vector<int> colptr, rowidx;       // CSC
    std::vector<T> val;
    SparseMatrix() : colptr(1, 0) {}
    int rows() const { return nr; }
    int cols() const { return nc; }
    // from triplets (duplicates are summed, as Matrix::sparseMatrix / TMB's tmbutils::asSparseMatrix do)
    static SparseMatrix from_triplets(int nr, int nc, long nnz, const int* ti, const int* tj, const double* tx) {
        SparseMatrix s;
        s.nr = nr; s.nc = nc;
        std::vector<long> start((size_t)nc + 1, 0);
        for (long k = 0; k < nnz; k++) start[(size_t)tj[k] + 1]++;
        for (int j = 0; j < nc; j++) start[(size_t)j + 1] += start[(size_t)j];
        std::vector<std::pair<int, double>> ent((size_t)nnz);
        std::vector<long> fill(start.begin(), start.end() - 1);
        for (long k = 0; k < nnz; k++) ent[(size_t)fill[(size_t)tj[k]]++] = std::make_pair(ti[k], tx[k]);
        s.colptr.assign((size_t)nc + 1, 0);
        for (int j = 0; j < nc; j++) {
            auto b = ent.begin() + start[(size_t)j], e = ent.begin() + start[(size_t)j + 1];
            std::stable_sort(b, e, [](const std::pair<int, double>& p, const std::pair<int, double>& q) { return p.first < q.first; });
            for (auto it = b; it != e; ++it) {
                if (it != b && (it - 1)->first == it->first) s.val.back() = s.val.back() + T(it->second);
                else { s.rowidx.push_back(it->first); s.val.push_back(T(it->second)); }
            }
            s.colptr[(size_t)j + 1] = (int)s.rowidx.size();
        }
        return s;
    }
    SparseMatrix block(int r0, int c0, int r, int c) const {
        SparseMatrix b;
        b.nr = r; b.nc = c;
        b.colptr.assign((size_t)c + 1, 0);
        for (int j = 0; j < c; j++) {
            for (int k = colptr[(size_t)(c0 + j)]; k < colptr[(size_t)(c0 + j) + 1]; k++)
                if (rowidx[(size_t)k] >= r0 && rowidx[(size_t)k] < r0 + r) { b.rowidx.push_back(rowidx[(size_t)k] - r0); b.val.push_back(val[(size_t)k]); }
            b.colptr[(size_t)j + 1] = (int)b.rowidx.size();
        }
        return b;
    }
    ::vector<T> col(int j) const {
        ::vector<T> v(nr);
        v.setZero();
        for (int k = colptr[(size_t)j]; k < colptr[(size_t)j + 1]; k++) v.x[(size_t)rowidx[(size_t)k]] = val[(size_t)k];
        return v;
    }
};
}  // namespace Eigen

template <class T> matrix<T>::matrix(const Eigen::SparseMatrix<T>& s) : x((size_t)s.nr * (size_t)s.nc, T(0.0)), nr(s.nr), nc(s.nc) {
    for (int j = 0; j < s.nc; j++)
        for (int k = s.colptr[(size_t)j]; k < s.colptr[(size_t)j + 1]; k++) (*this)(s.rowidx[(size_t)k], j) = s.val[(size_t)k];
}
// TMB: SparseMatrix<Type> * vector<Type> -> vector<Type>
template <class T> inline vector<T> operator*(const Eigen::SparseMatrix<T>& A, const vector<T>& v) {
    if (A.nc != v.size()) throw std::runtime_error("sparse * vector size mismatch");
    vector<T> r(A.nr);
    std::vector<char> touched((size_t)A.nr, 0);
    r.setZero();
    for (int j = 0; j < A.nc; j++)
        for (int k = A.colptr[(size_t)j]; k < A.colptr[(size_t)j + 1]; k++) {
            size_t i = (size_t)A.rowidx[(size_t)k];
            T t = A.val[(size_t)k] * v.x[(size_t)j];
            if (touched[i]) r.x[i] = r.x[i] + t; else { r.x[i] = t; touched[i] = 1; }
        }
    return r;
}
namespace tmbutils {
template <class T> inline Eigen::SparseMatrix<T> asSparseMatrix(const ::matrix<T>& m) {
    Eigen::SparseMatrix<T> s;
    s.nr = m.nr; s.nc = m.nc;
    s.colptr.assign((size_t)m.nc + 1, 0);
    for (int j = 0; j < m.nc; j++) {
        for (int i = 0; i < m.nr; i++)
            if (asDouble(m(i, j)) != 0.0) { s.rowidx.push_back(i); s.val.push_back(m(i, j)); }
        s.colptr[(size_t)j + 1] = (int)s.rowidx.size();
    }
    return s;
}
}  // namespace tmbutils
using tmbutils::asSparseMatrix;

// ------------------------------------------------------------------------------------------
// TMB's distribution / linear-algebra helpers
// ------------------------------------------------------------------------------------------
// TMB dnorm (distributions_R.hpp): -log(sqrt(2 pi) sd) - resid^2 / 2
template <class T> inline T dnorm(T x, T mean, T sd, int give_log = 0) {
    T resid = (x - mean) / sd;
    T logans = T(-std::log(std::sqrt(2 * M_PI))) - log(sd) - T(.5) * resid * resid;
    if (give_log) return logans;
    return exp(logans);
}
// TMB dt (distributions_R.hpp)
template <class T> inline T dt(T x, T df, int give_log) {
    T logres = lgamma((df + 1) / 2) - T(1) / 2 * log(df * M_PI) - lgamma(df / 2) - (df + 1) / 2 * log(1 + x * x / df);
    if (!give_log) return exp(logres);
    return logres;
}
inline double besselI(double x, double nu) { return ad::bessel_i(x, nu); }
using ad::besselI;

namespace atomic {
// log-determinant: LU with partial pivoting (TMB evaluates X.determinant() and takes the log)
template <class T> inline T logdet(::matrix<T> a) {
    int n = a.rows();
    T ld(0.0);
    int sign = 1;
    for (int c = 0; c < n; c++) {
        int p = c;
        double best = std::fabs(asDouble(a(c, c)));
        for (int r = c + 1; r < n; r++) { double v = std::fabs(asDouble(a(r, c))); if (v > best) { best = v; p = r; } }
        if (p != c) { for (int j = 0; j < n; j++) std::swap(a(c, j), a(p, j)); sign = -sign; }
        T piv = a(c, c);
        if (asDouble(piv) < 0) { sign = -sign; ld = ld + log(-piv); } else ld = ld + log(piv);
        for (int r = c + 1; r < n; r++) {
            T f = a(r, c) / piv;
            for (int j = c; j < n; j++) a(r, j) = a(r, j) - f * a(c, j);
        }
    }
    if (sign < 0) return T(std::nan(""));
    return ld;
}
// inverse of a positive definite matrix + its log-determinant (Cholesky)
template <class T> inline ::matrix<T> matinvpd(::matrix<T> a, T& logdet_out) {
    int n = a.rows();
    ::matrix<T> L(n, n);
    L.setZero();
    T ld(0.0);
    for (int j = 0; j < n; j++) {
        T s = a(j, j);
        for (int k = 0; k < j; k++) s = s - L(j, k) * L(j, k);
        T d = sqrt(s);
        L(j, j) = d;
        ld = ld + log(s);
        for (int i = j + 1; i < n; i++) {
            T t = a(i, j);
            for (int k = 0; k < j; k++) t = t - L(i, k) * L(j, k);
            L(i, j) = t / d;
        }
    }
    logdet_out = ld;
    return a.inverse();
}
}  // namespace atomic

namespace density {
template <class T>
struct GMRF_t {
    Eigen::SparseMatrix<T> Q;
    // x' Q x
    T Quadform(::vector<T> x) { return (x * (Q * x)).sum(); }
};
template <class T> inline GMRF_t<T> GMRF(const Eigen::SparseMatrix<T>& Q) { return GMRF_t<T>{Q}; }
}  // namespace density
namespace R_inla {}

// ------------------------------------------------------------------------------------------
// objective_function<Type>: data / parameter tables behind the DATA_* / PARAMETER* macros
// ------------------------------------------------------------------------------------------
struct shim_item {
    int kind = 0;                              // 0 double array, 1 int array, 2 string, 3 sparse triplets
    std::vector<long> dim;
    const double* d = nullptr;
    const int* i = nullptr;
    std::string s;
    const int* ti = nullptr;
    const int* tj = nullptr;
    long nnz = 0;
};
struct shim_report {
    std::vector<double> x;
    std::vector<long> dim;
};

template <class Type>
class objective_function {
public:
    const std::map<std::string, shim_item>* data = nullptr;
    std::map<std::string, std::vector<Type>> par;
    std::map<std::string, shim_report>* reports = nullptr;
    int current_parallel_region = -1;

    const shim_item& item(const char* name, int kind) const {
        auto it = data->find(name);
        if (it == data->end()) throw std::runtime_error(std::string("missing data item: ") + name);
        if (it->second.kind != kind) throw std::runtime_error(std::string("wrong kind for data item: ") + name);
        return it->second;
    }
    static long count(const shim_item& it) { long c = 1; for (long v : it.dim) c *= v; return c; }
    std::string data_string(const char* name) const { return item(name, 2).s; }
    int data_integer(const char* name) const { return item(name, 1).i[0]; }
    vector<Type> data_vector(const char* name) const {
        const shim_item& it = item(name, 0);
        long n = count(it);
        vector<Type> v(n);
        for (long k = 0; k < n; k++) v.x[(size_t)k] = Type(it.d[k]);
        return v;
    }
    vector<int> data_ivector(const char* name) const {
        const shim_item& it = item(name, 1);
        long n = count(it);
        vector<int> v(n);
        for (long k = 0; k < n; k++) v.x[(size_t)k] = it.i[k];
        return v;
    }
    matrix<Type> data_matrix(const char* name) const {
        const shim_item& it = item(name, 0);
        if (it.dim.size() != 2) throw std::runtime_error(std::string("not a matrix: ") + name);
        matrix<Type> m((int)it.dim[0], (int)it.dim[1]);
        for (size_t k = 0; k < m.x.size(); k++) m.x[k] = Type(it.d[k]);
        return m;
    }
    tmbutils::array<Type> data_array(const char* name) const {
        const shim_item& it = item(name, 0);
        tmbutils::array<Type> a;
        long n = count(it);
        a.x.resize((size_t)n);
        for (long k = 0; k < n; k++) a.x[(size_t)k] = Type(it.d[k]);
        for (long v : it.dim) a.dim.push_back((int)v);
        return a;
    }
    Eigen::SparseMatrix<Type> data_sparse(const char* name) const {
        const shim_item& it = item(name, 3);
        return Eigen::SparseMatrix<Type>::from_triplets((int)it.dim[0], (int)it.dim[1], it.nnz, it.ti, it.tj, it.d);
    }
    const std::vector<Type>& param(const char* name) const {
        auto it = par.find(name);
        if (it == par.end()) throw std::runtime_error(std::string("missing parameter: ") + name);
        return it->second;
    }
    Type param_scalar(const char* name) const { return param(name).at(0); }
    vector<Type> param_vector(const char* name) const { return vector<Type>(param(name)); }
    template <class M> void report(const char* name, const M& m) {
        if (!reports) return;
        shim_report r;
        r.x.resize(m.x.size());
        for (size_t k = 0; k < m.x.size(); k++) r.x[k] = asDouble(m.x[k]);
        r.dim = {m.rows(), m.cols()};
        (*reports)[name] = r;
    }
    Type operator()();
};

#define TMB_OBJECTIVE_PTR this
#define DATA_STRING(name) std::string name = TMB_OBJECTIVE_PTR->data_string(#name);
#define DATA_INTEGER(name) int name = TMB_OBJECTIVE_PTR->data_integer(#name);
#define DATA_VECTOR(name) vector<Type> name(TMB_OBJECTIVE_PTR->data_vector(#name));
#define DATA_IVECTOR(name) vector<int> name(TMB_OBJECTIVE_PTR->data_ivector(#name));
#define DATA_MATRIX(name) matrix<Type> name(TMB_OBJECTIVE_PTR->data_matrix(#name));
#define DATA_ARRAY(name) tmbutils::array<Type> name(TMB_OBJECTIVE_PTR->data_array(#name));
#define DATA_SPARSE_MATRIX(name) Eigen::SparseMatrix<Type> name(TMB_OBJECTIVE_PTR->data_sparse(#name));
#define PARAMETER(name) Type name(TMB_OBJECTIVE_PTR->param_scalar(#name));
#define PARAMETER_VECTOR(name) vector<Type> name(TMB_OBJECTIVE_PTR->param_vector(#name));
// TMB reports only from the double-typed evaluation
#define REPORT(name) if (isDouble<Type>::value && TMB_OBJECTIVE_PTR->current_parallel_region < 0) { TMB_OBJECTIVE_PTR->report(#name, name); }

#endif
