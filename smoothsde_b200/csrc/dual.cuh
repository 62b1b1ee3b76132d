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

}  // namespace ssde
