// Forward-mode (tangent) scalar for second derivatives.
//
// TMB gets the Hessian of the objective by taping the gradient tape again (MakeADHessObject2,
// src/init.c:13).  Here every piece of arithmetic on the hot path (linear predictor, natural-scale
// transform, filter step, scan combines, hand-derived adjoint, transposed design product, penalty)
// is a template over its scalar type R: R = double gives the nllk + gradient kernels, R = Dual runs
// the very same code on (value, directional derivative) pairs, so one forward + adjoint pass with
// theta = theta0 + eps * v returns the gradient AND the exact Hessian-vector product H v (tangent
// of the adjoint = second-order adjoint).  Compositions of tangent-augmented scan elements are
// again associative, so the time-parallel structure carries over unchanged.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define SSDE_HD __host__ __device__ __forceinline__
#else
#define SSDE_HD inline
#endif

namespace ssde {

struct Dual {
    double v, d;
    Dual() = default;
    SSDE_HD Dual(double v_) : v(v_), d(0.0) {}
    SSDE_HD Dual(double v_, double d_) : v(v_), d(d_) {}
};

// ---- value / tangent access that also works for plain doubles
SSDE_HD double value(double x) { return x; }
SSDE_HD double value(const Dual& x) { return x.v; }
SSDE_HD double tangent(double) { return 0.0; }
SSDE_HD double tangent(const Dual& x) { return x.d; }

template <class R> struct ScalarOf;
template <> struct ScalarOf<double> {
    static constexpr int NDBL = 1;
    static SSDE_HD double make(double v, double) { return v; }
};
template <> struct ScalarOf<Dual> {
    static constexpr int NDBL = 2;
    static SSDE_HD Dual make(double v, double d) { return Dual(v, d); }
};

// ---- arithmetic
SSDE_HD Dual operator-(const Dual& a) { return Dual(-a.v, -a.d); }
SSDE_HD Dual operator+(const Dual& a, const Dual& b) { return Dual(a.v + b.v, a.d + b.d); }
SSDE_HD Dual operator+(const Dual& a, double b) { return Dual(a.v + b, a.d); }
SSDE_HD Dual operator+(double a, const Dual& b) { return Dual(a + b.v, b.d); }
SSDE_HD Dual operator-(const Dual& a, const Dual& b) { return Dual(a.v - b.v, a.d - b.d); }
SSDE_HD Dual operator-(const Dual& a, double b) { return Dual(a.v - b, a.d); }
SSDE_HD Dual operator-(double a, const Dual& b) { return Dual(a - b.v, -b.d); }
SSDE_HD Dual operator*(const Dual& a, const Dual& b) { return Dual(a.v * b.v, fma(a.v, b.d, a.d * b.v)); }
SSDE_HD Dual operator*(const Dual& a, double b) { return Dual(a.v * b, a.d * b); }
SSDE_HD Dual operator*(double a, const Dual& b) { return Dual(a * b.v, a * b.d); }
SSDE_HD Dual operator/(const Dual& a, const Dual& b) {
    const double i = 1.0 / b.v, q = a.v * i;
    return Dual(q, (a.d - q * b.d) * i);
}
SSDE_HD Dual operator/(const Dual& a, double b) { const double i = 1.0 / b; return Dual(a.v * i, a.d * i); }
SSDE_HD Dual operator/(double a, const Dual& b) {
    const double i = 1.0 / b.v, q = a * i;
    return Dual(q, -q * b.d * i);
}
SSDE_HD Dual& operator+=(Dual& a, const Dual& b) { a.v += b.v; a.d += b.d; return a; }
SSDE_HD Dual& operator+=(Dual& a, double b) { a.v += b; return a; }
SSDE_HD Dual& operator-=(Dual& a, const Dual& b) { a.v -= b.v; a.d -= b.d; return a; }
SSDE_HD Dual& operator-=(Dual& a, double b) { a.v -= b; return a; }
SSDE_HD Dual& operator*=(Dual& a, const Dual& b) { a = a * b; return a; }
SSDE_HD Dual& operator*=(Dual& a, double b) { a.v *= b; a.d *= b; return a; }

using ::exp;
using ::log;
using ::sqrt;
using ::fma;
SSDE_HD Dual exp(const Dual& a) { const double e = ::exp(a.v); return Dual(e, e * a.d); }
SSDE_HD Dual log(const Dual& a) { return Dual(::log(a.v), a.d / a.v); }
SSDE_HD Dual sqrt(const Dual& a) { const double s = ::sqrt(a.v); return Dual(s, 0.5 * a.d / s); }

// c + a * b with a plain-double factor (design value times coefficient)
SSDE_HD double fmad(double a, double b, double c) { return fma(a, b, c); }
SSDE_HD Dual fmad(double a, const Dual& b, const Dual& c) { return Dual(fma(a, b.v, c.v), fma(a, b.d, c.d)); }

// ---------------------------------------------------------------------------------------------
// DualN<N>: value + N partial derivatives (forward mode along N seed directions at once).  Used by
// the one-pass data-term Hessian of the BM / OU models (kernels_sde.cuh: sde_hess_kernel): the
// closed-form d nllk_i / d eta of a row (sde_row) evaluated on DualN<NP> numbers seeded with the
// unit directions of the row's NP linear predictors yields the exact NP x NP block W_i.
// ---------------------------------------------------------------------------------------------
template <int N>
struct DualN {
    double v, d[N];
    DualN() = default;
    SSDE_HD DualN(double v_) : v(v_) {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = 0.0;
    }
};
template <int N> SSDE_HD double value(const DualN<N>& x) { return x.v; }
#define SSDE_DN_LOOP _Pragma("unroll") for (int i = 0; i < N; ++i)
template <int N> SSDE_HD DualN<N> operator-(const DualN<N>& a) { DualN<N> r; r.v = -a.v; SSDE_DN_LOOP r.d[i] = -a.d[i]; return r; }
template <int N> SSDE_HD DualN<N> operator+(const DualN<N>& a, const DualN<N>& b) { DualN<N> r; r.v = a.v + b.v; SSDE_DN_LOOP r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> SSDE_HD DualN<N> operator-(const DualN<N>& a, const DualN<N>& b) { DualN<N> r; r.v = a.v - b.v; SSDE_DN_LOOP r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> SSDE_HD DualN<N> operator*(const DualN<N>& a, const DualN<N>& b) { DualN<N> r; r.v = a.v * b.v; SSDE_DN_LOOP r.d[i] = fma(a.v, b.d[i], a.d[i] * b.v); return r; }
template <int N> SSDE_HD DualN<N> operator/(const DualN<N>& a, const DualN<N>& b) {
    DualN<N> r; const double ib = 1.0 / b.v; r.v = a.v * ib; SSDE_DN_LOOP r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r;
}
template <int N> SSDE_HD DualN<N> operator+(const DualN<N>& a, double b) { DualN<N> r = a; r.v += b; return r; }
template <int N> SSDE_HD DualN<N> operator+(double a, const DualN<N>& b) { DualN<N> r = b; r.v += a; return r; }
template <int N> SSDE_HD DualN<N> operator-(const DualN<N>& a, double b) { DualN<N> r = a; r.v -= b; return r; }
template <int N> SSDE_HD DualN<N> operator-(double a, const DualN<N>& b) { DualN<N> r; r.v = a - b.v; SSDE_DN_LOOP r.d[i] = -b.d[i]; return r; }
template <int N> SSDE_HD DualN<N> operator*(const DualN<N>& a, double b) { DualN<N> r; r.v = a.v * b; SSDE_DN_LOOP r.d[i] = a.d[i] * b; return r; }
template <int N> SSDE_HD DualN<N> operator*(double a, const DualN<N>& b) { return b * a; }
template <int N> SSDE_HD DualN<N> operator/(const DualN<N>& a, double b) { return a * (1.0 / b); }
template <int N> SSDE_HD DualN<N> operator/(double a, const DualN<N>& b) {
    DualN<N> r; const double ib = 1.0 / b.v; r.v = a * ib; SSDE_DN_LOOP r.d[i] = -r.v * b.d[i] * ib; return r;
}
template <int N> SSDE_HD DualN<N>& operator+=(DualN<N>& a, const DualN<N>& b) { a.v += b.v; SSDE_DN_LOOP a.d[i] += b.d[i]; return a; }
template <int N> SSDE_HD DualN<N>& operator+=(DualN<N>& a, double b) { a.v += b; return a; }
template <int N> SSDE_HD DualN<N>& operator-=(DualN<N>& a, const DualN<N>& b) { a.v -= b.v; SSDE_DN_LOOP a.d[i] -= b.d[i]; return a; }
template <int N> SSDE_HD DualN<N> exp(const DualN<N>& a) { DualN<N> r; r.v = ::exp(a.v); SSDE_DN_LOOP r.d[i] = r.v * a.d[i]; return r; }
template <int N> SSDE_HD DualN<N> log(const DualN<N>& a) { DualN<N> r; const double ia = 1.0 / a.v; r.v = ::log(a.v); SSDE_DN_LOOP r.d[i] = a.d[i] * ia; return r; }
template <int N> SSDE_HD DualN<N> sqrt(const DualN<N>& a) { DualN<N> r; r.v = ::sqrt(a.v); const double h = 0.5 / r.v; SSDE_DN_LOOP r.d[i] = a.d[i] * h; return r; }
#undef SSDE_DN_LOOP

}  // namespace ssde
