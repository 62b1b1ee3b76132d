// CTCRW Kalman filter algebra for the time-parallel (associative-scan) likelihood engine.
//
// Everything in this header is a small inline function usable from device code (the CUDA
// kernels in kernels_ctcrw.cuh) and from host code (tests/harness/math_harness.cpp compiles it
// with g++ to check the algebra against the oracle without a GPU -- test infrastructure only,
// the product path always runs these functions on the device).
//
// What is being computed (reference: src/nllk/nllk_ctcrw.hpp:195-247, helpers :30-91):
//   state at row i = predicted (a, P) for row i;  row i (same track, obs present) does
//     u = y - Z a;  F = Z P Z' + H;  llk -= (log|F| + u'F^-1 u)/2;
//     K = T P Z' F^-1;  a <- T a + K u + B mu;  P <- T P L' + Q,  L = T - K Z
//   with T, Q, B built from (beta_i, sigma_i, dt_i = t_{i+1} - t_i).
//
// Decoupled representation.  With H = sigma_obs^2 I (makeH_ctcrw, :30-38) and a block-diagonal
// P0 with identical 2x2 blocks (the default diag(1,10,1,10), R/sde.R:584) the 2d-state filter
// is d independent 2-state (position, velocity) filters that share one covariance recursion:
//   P = [[p11,p12],[p12,p22]] (3 doubles),  a_d = (z_d, v_d) per dimension,
//   F = p11 + h,  log|F| = d log F,  u'F^-1u = sum_d u_d^2 / F.
//
// Time-parallel formulation (Sarkka & Garcia-Fernandez 2021, prediction form).  Row i is the map
//   s_i -> s_{i+1}:  "update with y_i, then predict with (T_i, Q_i, c_i)",
// which is the element (A, b, C, eta, J) = (T_i, c_i, Q_i, Z'y_i/h, Z'Z/h); a missing row has
// eta = J = 0 and a track-start row is the constant map (A = 0, b = a0, C = P0).  Elements
// compose associatively (fwd_combine); a thread composes its consecutive rows with the cheap
// fwd_append (one Kalman step on (b, C) plus the A/eta/J bookkeeping).
//
// The hand-derived adjoint (gradient) is again a scan, in reverse time, over elements
// (L, z, D):  (abar, Pbar) <- (L'abar+ - z,  L'Pbar+ L + sum_d sym(L'abar+_d z_d') + D).
#pragma once

#include <math.h>

#include "dual.cuh"

namespace ssde {

// ---------------------------------------------------------------------------------------------
// small types
// ---------------------------------------------------------------------------------------------
// Every type and function below is a template over the scalar R (double, or Dual for the
// tangent-augmented pass that yields Hessian-vector products, dual.cuh); R defaults to double.
template <class R>
struct Sym2T {           // symmetric 2x2: [[a, b], [b, c]]
    R a, b, c;
};
template <class R>
struct Mat2T {           // general 2x2: [[m11, m12], [m21, m22]]
    R m11, m12, m21, m22;
};
template <class R>
struct Vec2T {
    R x, y;
};
using Sym2 = Sym2T<double>;
using Mat2 = Mat2T<double>;
using Vec2 = Vec2T<double>;

template <int ND, class R = double>
struct State {           // predicted state: per-dimension mean (z, v) + shared covariance
    Vec2T<R> a[ND];
    Sym2T<R> P;
};

// Per-row step quantities derived from the transformed parameters (tau, e = exp(-dt/tau),
// s2 = sigma^2) and dt.  makeT/makeQ/makeB_ctcrw, nllk_ctcrw.hpp:45-91, written in tau = 1/beta
// so that no division is needed:  (1-e)/beta = (1-e) tau, (sigma/beta)^2 = s2 tau^2, ...
template <class R>
struct StepParT {
    R T12, e;            // T = [[1, T12], [0, e]]
    Sym2T<R> Q;
    R B1, B2;            // B mu = (B1 mu, B2 mu)
};
using StepPar = StepParT<double>;

template <class R>
SSDE_HD StepParT<R> make_step(const R& tau, const R& e, const R& s2, double dt) {
    StepParT<R> s;
    const R om = 1.0 - e;
    const R ome2 = 1.0 - e * e;
    s.T12 = om * tau;
    s.e = e;
    const R st2 = s2 * tau * tau;
    s.Q.a = st2 * (dt - 2.0 * tau * om + 0.5 * tau * ome2);   // :68-69
    s.Q.b = 0.5 * st2 * om * om;                              // :70
    s.Q.c = 0.5 * s2 * tau * ome2;                            // :72
    s.B1 = dt - om * tau;                                     // :87
    s.B2 = om;                                                // :88
    return s;
}

// Natural-scale transform of one linear-predictor row, nllk_ctcrw.hpp:152-156:
//   tau = exp(eta_tau), nu = exp(eta_nu), sigma = 2 nu / sqrt(pi tau)  =>  s2 = 4 nu^2/(pi tau)
template <class R>
SSDE_HD void transform_row(const R& eta_tau, const R& eta_nu, double dt, R& tau, R& e, R& s2) {
    tau = exp(eta_tau);
    const R nu = exp(eta_nu);
    const R itau = 1.0 / tau;
    s2 = (4.0 / 3.14159265358979323846) * nu * nu * itau;
    e = exp(-dt * itau);
}

// ---------------------------------------------------------------------------------------------
// sequential filter step (prediction form), one row
// ---------------------------------------------------------------------------------------------
// Intermediate quantities of a step that the adjoint needs.
// keeps a parameter out of template-argument deduction (so that nullptr can be passed)
template <class T>
struct Ident {
    using type = T;
};

template <int ND, class R = double>
struct StepAux {
    R F, iF, g1, g2;             // F = p11 + h, gain G = (g1, g2) = P e1 / F
    R w[ND];                     // w_d = u_d / F
    R afv[ND];                   // filtered velocity a_f,d,v
    R tp12, tp22;                // (T P_f)_12, (T P_f)_22
};

// Advances `s` (predicted state of row i) to the predicted state of row i+1.  Outputs the
// innovation variance F (1 if the row is missing) and quad = sum_d u_d^2 / F (0 if missing); the
// row's log-likelihood contribution is  -(d log F + quad)/2  (:231-234, no d*log(2 pi)).
// `has_obs` false reproduces the missing branch (:214-217).
template <int ND, bool WITH_AUX, class R>
SSDE_HD void fwd_step_q(State<ND, R>& s, const StepParT<R>& sp, const double* y, const R* mu,
                        bool has_obs, const R& h, typename Ident<StepAux<ND, R>>::type* aux, R& F_out,
                        R& quad_out) {
    Sym2T<R> Pf = s.P;
    Vec2T<R> af[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) af[d] = s.a[d];
    F_out = 1.0;
    quad_out = 0.0;
    if (has_obs) {
        const R F = s.P.a + h;                     // :223
        const R iF = 1.0 / F;
        const R g1 = s.P.a * iF, g2 = s.P.b * iF;
        R quad = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const R u = y[d] - s.a[d].x;           // :221
            const R w = u * iF;
            quad += u * w;
            af[d].x = s.a[d].x + g1 * u;
            af[d].y = s.a[d].y + g2 * u;
            if (WITH_AUX) aux->w[d] = w;
        }
        F_out = F;
        quad_out = quad;
        Pf.a = s.P.a - g1 * s.P.a;                      // P - F G G'
        Pf.b = s.P.b - g1 * s.P.b;
        Pf.c = s.P.c - g2 * s.P.b;
        if (WITH_AUX) { aux->F = F; aux->iF = iF; aux->g1 = g1; aux->g2 = g2; }
    } else if (WITH_AUX) {
        aux->F = 1.0; aux->iF = 0.0; aux->g1 = 0.0; aux->g2 = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) aux->w[d] = 0.0;
    }
    // predict: a+ = T a_f + B mu,  P+ = T P_f T' + Q      (:238-241 / :216-217)
    const R tp11 = Pf.a + sp.T12 * Pf.b;
    const R tp12 = Pf.b + sp.T12 * Pf.c;
    const R tp22 = sp.e * Pf.c;
    if (WITH_AUX) { aux->tp12 = tp12; aux->tp22 = tp22; }
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        if (WITH_AUX) aux->afv[d] = af[d].y;
        s.a[d].x = af[d].x + sp.T12 * af[d].y + sp.B1 * mu[d];
        s.a[d].y = sp.e * af[d].y + sp.B2 * mu[d];
    }
    s.P.a = tp11 + sp.T12 * tp12 + sp.Q.a;
    s.P.b = sp.e * tp12 + sp.Q.b;
    s.P.c = sp.e * tp22 + sp.Q.c;
}

// Same step, returning the row's log-likelihood contribution.
template <int ND, bool WITH_AUX, class R>
SSDE_HD R fwd_step(State<ND, R>& s, const StepParT<R>& sp, const double* y, const R* mu,
                   bool has_obs, const R& h, typename Ident<StepAux<ND, R>>::type* aux) {
    R F, quad;
    fwd_step_q<ND, WITH_AUX, R>(s, sp, y, mu, has_obs, h, aux, F, quad);
    return has_obs ? R(-0.5 * ((double)ND * log(F) + quad)) : R(0.0);
}

// ---------------------------------------------------------------------------------------------
// forward scan elements
// ---------------------------------------------------------------------------------------------
template <int ND, class R = double>
struct FwdElem {
    Mat2T<R> A;
    Vec2T<R> b[ND];
    Sym2T<R> C;
    Vec2T<R> eta[ND];
    Sym2T<R> J;
    static constexpr int NDBL = (4 + 2 * ND + 3 + 2 * ND + 3) * ScalarOf<R>::NDBL;   // size in doubles
};

template <int ND, class R = double>
SSDE_HD FwdElem<ND, R> fwd_identity() {
    FwdElem<ND, R> E;
    E.A = {1.0, 0.0, 0.0, 1.0};
    E.C = {0.0, 0.0, 0.0};
    E.J = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < ND; ++d) { E.b[d] = {0.0, 0.0}; E.eta[d] = {0.0, 0.0}; }
    return E;
}

// Constant map to a known state (track start, or the incoming state of a time shard).
template <int ND, class R>
SSDE_HD FwdElem<ND, R> fwd_const(const State<ND, R>& s) {
    FwdElem<ND, R> E = fwd_identity<ND, R>();
    E.A = {0.0, 0.0, 0.0, 0.0};
    E.C = s.P;
#pragma unroll
    for (int d = 0; d < ND; ++d) E.b[d] = s.a[d];
    return E;
}

// E <- (element of one ordinary row) o E.   Equivalent to fwd_combine(E, elem_row) but uses the
// structure J_row = e1 e1'/h, eta_row = e1 y/h:  (b, C) advance by one Kalman step, A <- L A,
// eta += A_row1' w, J += A_row1' A_row1 / F.
template <int ND, class R>
SSDE_HD void fwd_append(FwdElem<ND, R>& E, const StepParT<R>& sp, const double* y, const R* mu,
                        bool has_obs, const R& h) {
    R L11 = 1.0, L21 = 0.0;                // L = T (I - G e1')
    if (has_obs) {
        const R F = E.C.a + h;
        const R iF = 1.0 / F;
        const R g1 = E.C.a * iF, g2 = E.C.b * iF;
        const R a11 = E.A.m11, a12 = E.A.m12;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const R u = y[d] - E.b[d].x;
            const R w = u * iF;
            E.eta[d].x += a11 * w;
            E.eta[d].y += a12 * w;
            E.b[d].x += g1 * u;
            E.b[d].y += g2 * u;
        }
        E.J.a += a11 * a11 * iF;
        E.J.b += a11 * a12 * iF;
        E.J.c += a12 * a12 * iF;
        const R c11 = E.C.a, c12 = E.C.b;
        E.C.a = c11 - g1 * c11;
        E.C.b = c12 - g1 * c12;
        E.C.c = E.C.c - g2 * c12;
        L11 = 1.0 - g1 - sp.T12 * g2;
        L21 = -sp.e * g2;
    }
    // A <- L A with L = [[L11, T12], [L21, e]]
    const Mat2T<R> A = E.A;
    E.A.m11 = L11 * A.m11 + sp.T12 * A.m21;
    E.A.m12 = L11 * A.m12 + sp.T12 * A.m22;
    E.A.m21 = L21 * A.m11 + sp.e * A.m21;
    E.A.m22 = L21 * A.m12 + sp.e * A.m22;
    // predict (b, C)
    const R tp11 = E.C.a + sp.T12 * E.C.b;
    const R tp12 = E.C.b + sp.T12 * E.C.c;
    const R tp22 = sp.e * E.C.c;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const R bx = E.b[d].x, by = E.b[d].y;
        E.b[d].x = bx + sp.T12 * by + sp.B1 * mu[d];
        E.b[d].y = sp.e * by + sp.B2 * mu[d];
    }
    E.C.a = tp11 + sp.T12 * tp12 + sp.Q.a;
    E.C.b = sp.e * tp12 + sp.Q.b;
    E.C.c = sp.e * tp22 + sp.Q.c;
}

// E <- (track-start element with state s0) o E : the constant map wins; eta, J keep describing
// the dependence of everything *before* the start on the incoming state.
template <int ND, class R>
SSDE_HD void fwd_append_start(FwdElem<ND, R>& E, const State<ND, R>& s0) {
    E.A = {0.0, 0.0, 0.0, 0.0};
    E.C = s0.P;
#pragma unroll
    for (int d = 0; d < ND; ++d) E.b[d] = s0.a[d];
}

// General composition: `Ei` covers earlier rows, `Ej` later rows.
//   M = (I + C_i J_j)^-1
//   A = A_j M A_i;  b = A_j M (b_i + C_i eta_j) + b_j;  C = A_j M C_i A_j' + C_j
//   eta = A_i' M' (eta_j - J_j b_i) + eta_i;  J = A_i' M' J_j A_i + J_i
template <int ND, class R>
SSDE_HD FwdElem<ND, R> fwd_combine(const FwdElem<ND, R>& Ei, const FwdElem<ND, R>& Ej) {
    FwdElem<ND, R> Ro;
    const Sym2T<R> C = Ei.C, J = Ej.J;
    const R x11 = 1.0 + C.a * J.a + C.b * J.b;
    const R x12 = C.a * J.b + C.b * J.c;
    const R x21 = C.b * J.a + C.c * J.b;
    const R x22 = 1.0 + C.b * J.b + C.c * J.c;
    const R idet = 1.0 / (x11 * x22 - x12 * x21);
    const Mat2T<R> M = {x22 * idet, -x12 * idet, -x21 * idet, x11 * idet};
    // AM = A_j M
    const Mat2T<R> Aj = Ej.A;
    const Mat2T<R> AM = {Aj.m11 * M.m11 + Aj.m12 * M.m21, Aj.m11 * M.m12 + Aj.m12 * M.m22,
                     Aj.m21 * M.m11 + Aj.m22 * M.m21, Aj.m21 * M.m12 + Aj.m22 * M.m22};
    const Mat2T<R> Ai = Ei.A;
    Ro.A = {AM.m11 * Ai.m11 + AM.m12 * Ai.m21, AM.m11 * Ai.m12 + AM.m12 * Ai.m22,
           AM.m21 * Ai.m11 + AM.m22 * Ai.m21, AM.m21 * Ai.m12 + AM.m22 * Ai.m22};
    // AMC = A_j M C_i  (2x2), then C = AMC A_j' + C_j
    const Mat2T<R> AMC = {AM.m11 * C.a + AM.m12 * C.b, AM.m11 * C.b + AM.m12 * C.c,
                      AM.m21 * C.a + AM.m22 * C.b, AM.m21 * C.b + AM.m22 * C.c};
    Ro.C.a = AMC.m11 * Aj.m11 + AMC.m12 * Aj.m12 + Ej.C.a;
    Ro.C.b = 0.5 * ((AMC.m11 * Aj.m21 + AMC.m12 * Aj.m22) + (AMC.m21 * Aj.m11 + AMC.m22 * Aj.m12))
            + Ej.C.b;
    Ro.C.c = AMC.m21 * Aj.m21 + AMC.m22 * Aj.m22 + Ej.C.c;
    // MtJ = M' J_j ;  J = A_i' MtJ A_i + J_i
    const Mat2T<R> MtJ = {M.m11 * J.a + M.m21 * J.b, M.m11 * J.b + M.m21 * J.c,
                      M.m12 * J.a + M.m22 * J.b, M.m12 * J.b + M.m22 * J.c};
    // AtN = A_i' M'  (2x2)
    const Mat2T<R> AtN = {Ai.m11 * M.m11 + Ai.m21 * M.m12, Ai.m11 * M.m21 + Ai.m21 * M.m22,
                      Ai.m12 * M.m11 + Ai.m22 * M.m12, Ai.m12 * M.m21 + Ai.m22 * M.m22};
    // W = A_i' (M' J_j)  then J = W A_i + J_i
    const Mat2T<R> W = {Ai.m11 * MtJ.m11 + Ai.m21 * MtJ.m21, Ai.m11 * MtJ.m12 + Ai.m21 * MtJ.m22,
                    Ai.m12 * MtJ.m11 + Ai.m22 * MtJ.m21, Ai.m12 * MtJ.m12 + Ai.m22 * MtJ.m22};
    Ro.J.a = W.m11 * Ai.m11 + W.m12 * Ai.m21 + Ei.J.a;
    Ro.J.b = 0.5 * ((W.m11 * Ai.m12 + W.m12 * Ai.m22) + (W.m21 * Ai.m11 + W.m22 * Ai.m21)) + Ei.J.b;
    Ro.J.c = W.m21 * Ai.m12 + W.m22 * Ai.m22 + Ei.J.c;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const Vec2T<R> bi = Ei.b[d], ej = Ej.eta[d];
        // t = b_i + C_i eta_j
        const R t1 = bi.x + C.a * ej.x + C.b * ej.y;
        const R t2 = bi.y + C.b * ej.x + C.c * ej.y;
        Ro.b[d].x = AM.m11 * t1 + AM.m12 * t2 + Ej.b[d].x;
        Ro.b[d].y = AM.m21 * t1 + AM.m22 * t2 + Ej.b[d].y;
        // r = eta_j - J_j b_i
        const R r1 = ej.x - (J.a * bi.x + J.b * bi.y);
        const R r2 = ej.y - (J.b * bi.x + J.c * bi.y);
        Ro.eta[d].x = AtN.m11 * r1 + AtN.m12 * r2 + Ei.eta[d].x;
        Ro.eta[d].y = AtN.m21 * r1 + AtN.m22 * r2 + Ei.eta[d].y;
    }
    return Ro;
}

// State after pushing state `s` through element E:  the (b, C) part of fwd_combine(const(s), E).
template <int ND, class R>
SSDE_HD State<ND, R> fwd_apply(const FwdElem<ND, R>& E, const State<ND, R>& s) {
    State<ND, R> r;
    const Sym2T<R> C = s.P, J = E.J;
    const R x11 = 1.0 + C.a * J.a + C.b * J.b;
    const R x12 = C.a * J.b + C.b * J.c;
    const R x21 = C.b * J.a + C.c * J.b;
    const R x22 = 1.0 + C.b * J.b + C.c * J.c;
    const R idet = 1.0 / (x11 * x22 - x12 * x21);
    const Mat2T<R> M = {x22 * idet, -x12 * idet, -x21 * idet, x11 * idet};
    const Mat2T<R> Aj = E.A;
    const Mat2T<R> AM = {Aj.m11 * M.m11 + Aj.m12 * M.m21, Aj.m11 * M.m12 + Aj.m12 * M.m22,
                     Aj.m21 * M.m11 + Aj.m22 * M.m21, Aj.m21 * M.m12 + Aj.m22 * M.m22};
    const Mat2T<R> AMC = {AM.m11 * C.a + AM.m12 * C.b, AM.m11 * C.b + AM.m12 * C.c,
                      AM.m21 * C.a + AM.m22 * C.b, AM.m21 * C.b + AM.m22 * C.c};
    r.P.a = AMC.m11 * Aj.m11 + AMC.m12 * Aj.m12 + E.C.a;
    r.P.b = 0.5 * ((AMC.m11 * Aj.m21 + AMC.m12 * Aj.m22) + (AMC.m21 * Aj.m11 + AMC.m22 * Aj.m12))
            + E.C.b;
    r.P.c = AMC.m21 * Aj.m21 + AMC.m22 * Aj.m22 + E.C.c;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const R t1 = s.a[d].x + C.a * E.eta[d].x + C.b * E.eta[d].y;
        const R t2 = s.a[d].y + C.b * E.eta[d].x + C.c * E.eta[d].y;
        r.a[d].x = AM.m11 * t1 + AM.m12 * t2 + E.b[d].x;
        r.a[d].y = AM.m21 * t1 + AM.m22 * t2 + E.b[d].y;
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// adjoint (reverse-time) elements
// ---------------------------------------------------------------------------------------------
// Adjoint of a predicted state: abar_d (2 per dim) and the symmetric FULL-matrix adjoint Pbar
// (inner product <Pbar, dP> = Pbar.a dp11 + 2 Pbar.b dp12 + Pbar.c dp22).
template <int ND, class R = double>
struct Adj {
    Vec2T<R> a[ND];
    Sym2T<R> P;
};

template <int ND, class R = double>
SSDE_HD Adj<ND, R> adj_zero() {
    Adj<ND, R> r;
    r.P = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < ND; ++d) r.a[d] = {0.0, 0.0};
    return r;
}

// Backward element:  abar_d = L' abar+_d - z_d ;
//                    Pbar   = L' Pbar+ L + sum_d sym(L' abar+_d z_d') + D
template <int ND, class R = double>
struct BwdElem {
    Mat2T<R> L;
    Vec2T<R> z[ND];
    Sym2T<R> D;
    static constexpr int NDBL = (4 + 2 * ND + 3) * ScalarOf<R>::NDBL;   // size in doubles
};

template <int ND, class R = double>
SSDE_HD BwdElem<ND, R> bwd_identity() {
    BwdElem<ND, R> E;
    E.L = {1.0, 0.0, 0.0, 1.0};
    E.D = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < ND; ++d) E.z[d] = {0.0, 0.0};
    return E;
}

// Constant map to a known adjoint (end of a track: zero; or the incoming adjoint of a shard).
template <int ND, class R>
SSDE_HD BwdElem<ND, R> bwd_const(const Adj<ND, R>& g) {
    BwdElem<ND, R> E;
    E.L = {0.0, 0.0, 0.0, 0.0};
    E.D = g.P;
#pragma unroll
    for (int d = 0; d < ND; ++d) E.z[d] = {-g.a[d].x, -g.a[d].y};
    return E;
}

// Elementary backward element of one ordinary row from its forward intermediates.
//   obs:      L = T (I - G e1'),  z_d = w_d e1,  D = Fl e1 e1',  Fl = (d/F - sum w^2)/2
//   missing:  L = T, z = 0, D = 0
//   `cut`:    the row is the last of its track: its prediction is discarded, i.e. the incoming
//             adjoint is replaced by zero before the row's own update terms are applied (L = 0).
template <int ND, class R>
SSDE_HD BwdElem<ND, R> bwd_row_elem(const StepParT<R>& sp, const StepAux<ND, R>& ax, bool has_obs, bool cut) {
    BwdElem<ND, R> E;
    R sw2 = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        E.z[d] = {has_obs ? ax.w[d] : 0.0, 0.0};
        sw2 += ax.w[d] * ax.w[d];
    }
    E.D = {has_obs ? 0.5 * ((double)ND * ax.iF - sw2) : 0.0, 0.0, 0.0};
    if (cut) {
        E.L = {0.0, 0.0, 0.0, 0.0};
    } else {
        E.L = {1.0 - ax.g1 - sp.T12 * ax.g2, sp.T12, -sp.e * ax.g2, sp.e};
    }
    return E;
}

// Composition: E1 covers EARLIER rows (applied last in reverse time), E2 later rows.
//   L = L2 L1;  z = L1' z2 + z1;  D = L1' D2 L1 + D1 - sum_d sym(L1' z2_d z1_d')
template <int ND, class R>
SSDE_HD BwdElem<ND, R> bwd_combine(const BwdElem<ND, R>& E1, const BwdElem<ND, R>& E2) {
    BwdElem<ND, R> Ro;
    const Mat2T<R> L1 = E1.L, L2 = E2.L;
    Ro.L = {L2.m11 * L1.m11 + L2.m12 * L1.m21, L2.m11 * L1.m12 + L2.m12 * L1.m22,
           L2.m21 * L1.m11 + L2.m22 * L1.m21, L2.m21 * L1.m12 + L2.m22 * L1.m22};
    // L1' D2 L1
    const Sym2T<R> D2 = E2.D;
    const R q11 = D2.a * L1.m11 + D2.b * L1.m21, q12 = D2.a * L1.m12 + D2.b * L1.m22;
    const R q21 = D2.b * L1.m11 + D2.c * L1.m21, q22 = D2.b * L1.m12 + D2.c * L1.m22;
    Ro.D.a = L1.m11 * q11 + L1.m21 * q21 + E1.D.a;
    Ro.D.b = 0.5 * ((L1.m11 * q12 + L1.m21 * q22) + (L1.m12 * q11 + L1.m22 * q21)) + E1.D.b;
    Ro.D.c = L1.m12 * q12 + L1.m22 * q22 + E1.D.c;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const R t1 = L1.m11 * E2.z[d].x + L1.m21 * E2.z[d].y;   // L1' z2
        const R t2 = L1.m12 * E2.z[d].x + L1.m22 * E2.z[d].y;
        const Vec2T<R> z1 = E1.z[d];
        Ro.z[d] = {t1 + z1.x, t2 + z1.y};
        Ro.D.a -= t1 * z1.x;
        Ro.D.b -= 0.5 * (t1 * z1.y + t2 * z1.x);
        Ro.D.c -= t2 * z1.y;
    }
    return Ro;
}

// bwd_combine(E1, bwd_row_elem(...)) for one ordinary row, without the products that the row
// element's structural zeros (z_d = (w_d, 0), D = diag(Fl, 0)) would only multiply by 0.
template <int ND, class R>
SSDE_HD BwdElem<ND, R> bwd_append_row(const BwdElem<ND, R>& E1, const StepParT<R>& sp, const StepAux<ND, R>& ax,
                                      bool has_obs, bool cut) {
    BwdElem<ND, R> Ro;
    const Mat2T<R> L1 = E1.L;
    if (cut) {
        Ro.L = {0.0, 0.0, 0.0, 0.0};
    } else {
        const R l11 = 1.0 - ax.g1 - sp.T12 * ax.g2, l21 = -sp.e * ax.g2;       // L2 = [[l11, T12], [l21, e]]
        Ro.L = {l11 * L1.m11 + sp.T12 * L1.m21, l11 * L1.m12 + sp.T12 * L1.m22,
                l21 * L1.m11 + sp.e * L1.m21, l21 * L1.m12 + sp.e * L1.m22};
    }
    R sw2 = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) sw2 += ax.w[d] * ax.w[d];
    const R Fl = has_obs ? R(0.5 * ((double)ND * ax.iF - sw2)) : R(0.0);
    const R q11 = Fl * L1.m11, q12 = Fl * L1.m12;
    Ro.D.a = L1.m11 * q11 + E1.D.a;
    Ro.D.b = 0.5 * (L1.m11 * q12 + L1.m12 * q11) + E1.D.b;
    Ro.D.c = L1.m12 * q12 + E1.D.c;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const R w = has_obs ? ax.w[d] : R(0.0);
        const R t1 = L1.m11 * w, t2 = L1.m12 * w;               // L1' z2 with z2 = (w, 0)
        const Vec2T<R> z1 = E1.z[d];
        Ro.z[d] = {t1 + z1.x, t2 + z1.y};
        Ro.D.a -= t1 * z1.x;
        Ro.D.b -= 0.5 * (t1 * z1.y + t2 * z1.x);
        Ro.D.c -= t2 * z1.y;
    }
    return Ro;
}

// bwd_apply(bwd_row_elem(...), g) for one ordinary row, again without the structural zeros.
template <int ND, class R>
SSDE_HD Adj<ND, R> bwd_apply_row(const StepParT<R>& sp, const StepAux<ND, R>& ax, bool has_obs, bool cut,
                                 const Adj<ND, R>& g) {
    Adj<ND, R> r;
    R sw2 = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) sw2 += ax.w[d] * ax.w[d];
    const R Fl = has_obs ? R(0.5 * ((double)ND * ax.iF - sw2)) : R(0.0);
    if (cut) {                                                   // L = 0: only the row's own terms survive
        r.P = {Fl, 0.0, 0.0};
#pragma unroll
        for (int d = 0; d < ND; ++d) r.a[d] = {has_obs ? R(-ax.w[d]) : R(0.0), 0.0};
        return r;
    }
    const Mat2T<R> L = {1.0 - ax.g1 - sp.T12 * ax.g2, sp.T12, -sp.e * ax.g2, sp.e};
    const Sym2T<R> P = g.P;
    const R q11 = P.a * L.m11 + P.b * L.m21, q12 = P.a * L.m12 + P.b * L.m22;
    const R q21 = P.b * L.m11 + P.c * L.m21, q22 = P.b * L.m12 + P.c * L.m22;
    r.P.a = L.m11 * q11 + L.m21 * q21 + Fl;
    r.P.b = 0.5 * ((L.m11 * q12 + L.m21 * q22) + (L.m12 * q11 + L.m22 * q21));
    r.P.c = L.m12 * q12 + L.m22 * q22;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const R w = has_obs ? ax.w[d] : R(0.0);
        const R t1 = L.m11 * g.a[d].x + L.m21 * g.a[d].y;       // L' abar+
        const R t2 = L.m12 * g.a[d].x + L.m22 * g.a[d].y;
        r.a[d] = {t1 - w, t2};
        r.P.a += t1 * w;
        r.P.b += 0.5 * (t2 * w);
    }
    return r;
}

// Adjoint at the start of E's range given the adjoint `g` flowing in at its end.
template <int ND, class R>
SSDE_HD Adj<ND, R> bwd_apply(const BwdElem<ND, R>& E, const Adj<ND, R>& g) {
    Adj<ND, R> r;
    const Mat2T<R> L = E.L;
    const Sym2T<R> P = g.P;
    const R q11 = P.a * L.m11 + P.b * L.m21, q12 = P.a * L.m12 + P.b * L.m22;
    const R q21 = P.b * L.m11 + P.c * L.m21, q22 = P.b * L.m12 + P.c * L.m22;
    r.P.a = L.m11 * q11 + L.m21 * q21 + E.D.a;
    r.P.b = 0.5 * ((L.m11 * q12 + L.m21 * q22) + (L.m12 * q11 + L.m22 * q21)) + E.D.b;
    r.P.c = L.m12 * q12 + L.m22 * q22 + E.D.c;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const R t1 = L.m11 * g.a[d].x + L.m21 * g.a[d].y;       // L' abar+
        const R t2 = L.m12 * g.a[d].x + L.m22 * g.a[d].y;
        r.a[d] = {t1 - E.z[d].x, t2 - E.z[d].y};
        r.P.a += t1 * E.z[d].x;
        r.P.b += 0.5 * (t1 * E.z[d].y + t2 * E.z[d].x);
        r.P.c += t2 * E.z[d].y;
    }
    return r;
}

// Chain rule from the adjoints of the step quantities (T12, T22 = e, B1, B2, Q11, Q12, Q22; Q12b
// counts both off-diagonal entries) to the linear predictors eta_tau, eta_nu.
template <class R>
SSDE_HD void ctcrw_chain(const R& T12b, const R& T22b, const R& B1b, const R& B2b, const R& Q11b, const R& Q12b,
                         const R& Q22b, const R& tau, const R& e, const R& s2, double dt, R& g_tau, R& g_nu) {
    const R om = 1.0 - e, ome2 = 1.0 - e * e;
    const R q1 = dt - 2.0 * tau * om + 0.5 * tau * ome2;
    const R st = s2 * tau;
    const R taub = (T12b - B1b) * om
                        + Q11b * (2.0 * st * q1 + st * tau * (0.5 * ome2 - 2.0 * om))
                        + Q12b * (st * om * om)
                        + Q22b * (0.5 * s2 * ome2);
    const R eb = (B1b - T12b) * tau + T22b - B2b
                      + Q11b * (st * tau * tau * (2.0 - e))
                      - Q12b * (st * tau * om)
                      - Q22b * (st * e);
    const R s2b = Q11b * (tau * tau * q1) + Q12b * (0.5 * tau * tau * om * om)
                       + Q22b * (0.5 * tau * ome2);
    g_tau = taub * tau + eb * e * dt / tau - s2b * s2;
    g_nu = 2.0 * s2b * s2;
}

// Gradient of one row w.r.t. its linear predictors, given the adjoint `g` of the state this row
// PREDICTS (row i+1's predicted state), the row's transformed parameters and its forward
// intermediates.  Outputs: gmu[d] = d nllk / d eta_mu_d,  g_tau = d/d eta_tau,  g_nu = d/d eta_nu,
// and the contribution to d nllk / d h  (h = sigma_obs^2), which also needs the update part.
template <int ND, class R>
SSDE_HD void row_param_grad(const Adj<ND, R>& g, const StepParT<R>& sp, const StepAux<ND, R>& ax,
                            const R* mu, const R& tau, const R& e, const R& s2, double dt,
                            bool has_obs, R* gmu, R& g_tau, R& g_nu, R& g_h) {
    // predict part: c_d = B mu_d, Q, T
    R B1b = 0.0, B2b = 0.0, T12b = 0.0, T22b = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        gmu[d] = sp.B1 * g.a[d].x + sp.B2 * g.a[d].y;
        B1b += g.a[d].x * mu[d];
        B2b += g.a[d].y * mu[d];
        T12b += g.a[d].x * ax.afv[d];
        T22b += g.a[d].y * ax.afv[d];
    }
    T12b += 2.0 * (g.P.a * ax.tp12 + g.P.b * ax.tp22);
    T22b += 2.0 * (g.P.b * ax.tp12 + g.P.c * ax.tp22);
    const R Q11b = g.P.a, Q12b = 2.0 * g.P.b, Q22b = g.P.c;
    ctcrw_chain(T12b, T22b, B1b, B2b, Q11b, Q12b, Q22b, tau, e, s2, dt, g_tau, g_nu);
    // update part (only h): hbar = G' Pf_bar G - sum_d w_d (af_bar_d' G) + Fl,
    // with af_bar_d = T' abar+_d and Pf_bar = T' Pbar+ T.
    g_h = 0.0;
    if (has_obs) {
        // T' Pbar T, T = [[1, T12], [0, e]]
        const R r11 = g.P.a;
        const R r12 = g.P.a * sp.T12 + g.P.b * sp.e;
        const R r22 = sp.T12 * (g.P.a * sp.T12 + g.P.b * sp.e)
                           + sp.e * (g.P.b * sp.T12 + g.P.c * sp.e);
        R sw2 = 0.0, acc = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const R f1 = g.a[d].x;                              // T' abar+
            const R f2 = sp.T12 * g.a[d].x + sp.e * g.a[d].y;
            acc += ax.w[d] * (f1 * ax.g1 + f2 * ax.g2);
            sw2 += ax.w[d] * ax.w[d];
        }
        g_h = ax.g1 * (r11 * ax.g1 + r12 * ax.g2) + ax.g2 * (r12 * ax.g1 + r22 * ax.g2) - acc
              + 0.5 * ((double)ND * ax.iF - sw2);
    }
}

}  // namespace ssde
