// Kalman filter algebra of the ONE-state-per-dimension state-space models of smoothSDE:
//   OU_SSM  src/nllk/nllk_ou_ssm.hpp:73-249  (makeT/B/Q_ou_ssm :30-69)
//   BM_SSM  src/nllk/nllk_bm_ssm.hpp:40-211  (makeQ_bm_ssm :27-36, drift = mu * dt :140)
// in the same time-parallel form as ctcrw_math.cuh (read that header first).
//
// Both templates run the loop of nllk_ctcrw with an n_dim-state filter whose T, B, Q are
// diagonal with EQUAL entries, Z = I and H = sigma_obs^2 I (makeH_*_ssm).  With P0 = c I (the
// default diag(rep(10, n_dim)), R/sde.R:553) the filter is n_dim independent SCALAR filters that
// share one variance recursion:
//     state at row i = predicted (a_d, p);   row i (obs present):
//     u_d = y_d - a_d;  F = p + h;  llk -= (d log F + sum_d u_d^2 / F) / 2;
//     g = p / F;  a_d <- t (a_d + g u_d) + c_d;  p <- t^2 (p - g p) + q
// with  OU_SSM: t = exp(-dt/tau), c_d = (1 - t) mu_d, q = kappa (1 - exp(-2 dt/tau))
//       BM_SSM: t = 1,            c_d = mu_d dt,      q = sigma^2 dt.
// Scan element of a row (prediction form): (A, b_d, C, eta_d, J) = (t, c_d, q, y_d/h, 1/h), all
// scalars; adjoint element (L, z_d, D) of  (abar, pbar) <- (L abar+ - z,  L^2 pbar+ + sum_d L abar+_d z_d + D).
#pragma once

#include "dual.cuh"

namespace ssde {

template <int ND, class R = double>
struct State1 {
    R a[ND];
    R p;
};

template <class R>
struct Step1 {
    R t, q;              // T = t I, Q = q I
    R cm;                // c_d = cm * mu_d   (B = cm I)
};

template <int ND, class R = double>
struct StepAux1 {
    R F, iF, g;
    R w[ND];             // u_d / F
    R af[ND];            // filtered means
    R pf;                // filtered variance
};

template <class T>
struct Ident1 {
    using type = T;
};

// one row: predicted state of row i -> predicted state of row i + 1
template <int ND, bool WITH_AUX, class R>
SSDE_HD void fwd_step1(State1<ND, R>& s, const Step1<R>& sp, const double* y, const R* mu, bool has_obs,
                       const R& h, typename Ident1<StepAux1<ND, R>>::type* aux, R& F_out, R& quad_out) {
    R pf = s.p;
    R af[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) af[d] = s.a[d];
    F_out = 1.0;
    quad_out = 0.0;
    if (has_obs) {
        const R F = s.p + h;
        const R iF = 1.0 / F;
        const R g = s.p * iF;
        R quad = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const R u = y[d] - s.a[d];
            const R w = u * iF;
            quad += u * w;
            af[d] = s.a[d] + g * u;
            if (WITH_AUX) aux->w[d] = w;
        }
        F_out = F;
        quad_out = quad;
        pf = s.p - g * s.p;
        if (WITH_AUX) { aux->F = F; aux->iF = iF; aux->g = g; }
    } else if (WITH_AUX) {
        aux->F = 1.0; aux->iF = 0.0; aux->g = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) aux->w[d] = 0.0;
    }
    if (WITH_AUX) aux->pf = pf;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        if (WITH_AUX) aux->af[d] = af[d];
        s.a[d] = sp.t * af[d] + sp.cm * mu[d];
    }
    s.p = sp.t * sp.t * pf + sp.q;
}

// ---- forward scan elements
template <int ND, class R = double>
struct FwdElem1 {
    R A;
    R b[ND];
    R C;
    R eta[ND];
    R J;
    static constexpr int NDBL = (3 + 2 * ND) * ScalarOf<R>::NDBL;
};

template <int ND, class R = double>
SSDE_HD FwdElem1<ND, R> fwd_identity1() {
    FwdElem1<ND, R> E;
    E.A = 1.0; E.C = 0.0; E.J = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) { E.b[d] = 0.0; E.eta[d] = 0.0; }
    return E;
}

// E <- (row) o E
template <int ND, class R>
SSDE_HD void fwd_append1(FwdElem1<ND, R>& E, const Step1<R>& sp, const double* y, const R* mu, bool has_obs, const R& h) {
    R L = sp.t;
    if (has_obs) {
        const R F = E.C + h;
        const R iF = 1.0 / F;
        const R g = E.C * iF;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const R u = y[d] - E.b[d];
            E.eta[d] += E.A * (u * iF);
            E.b[d] += g * u;
        }
        E.J += E.A * E.A * iF;
        E.C = E.C - g * E.C;
        L = sp.t * (1.0 - g);
    }
    E.A = L * E.A;
#pragma unroll
    for (int d = 0; d < ND; ++d) E.b[d] = sp.t * E.b[d] + sp.cm * mu[d];
    E.C = sp.t * sp.t * E.C + sp.q;
}

template <int ND, class R>
SSDE_HD void fwd_append_start1(FwdElem1<ND, R>& E, const State1<ND, R>& s0) {
    E.A = 0.0;
    E.C = s0.p;
#pragma unroll
    for (int d = 0; d < ND; ++d) E.b[d] = s0.a[d];
}

// Ei earlier rows, Ej later rows
template <int ND, class R>
SSDE_HD FwdElem1<ND, R> fwd_combine1(const FwdElem1<ND, R>& Ei, const FwdElem1<ND, R>& Ej) {
    FwdElem1<ND, R> Ro;
    const R M = 1.0 / (1.0 + Ei.C * Ej.J);
    const R AM = Ej.A * M;
    Ro.A = AM * Ei.A;
    Ro.C = AM * Ei.C * Ej.A + Ej.C;
    const R AiM = Ei.A * M;
    Ro.J = AiM * Ej.J * Ei.A + Ei.J;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        Ro.b[d] = AM * (Ei.b[d] + Ei.C * Ej.eta[d]) + Ej.b[d];
        Ro.eta[d] = AiM * (Ej.eta[d] - Ej.J * Ei.b[d]) + Ei.eta[d];
    }
    return Ro;
}

template <int ND, class R>
SSDE_HD State1<ND, R> fwd_apply1(const FwdElem1<ND, R>& E, const State1<ND, R>& s) {
    State1<ND, R> r;
    const R M = 1.0 / (1.0 + s.p * E.J);
    const R AM = E.A * M;
    r.p = AM * s.p * E.A + E.C;
#pragma unroll
    for (int d = 0; d < ND; ++d) r.a[d] = AM * (s.a[d] + s.p * E.eta[d]) + E.b[d];
    return r;
}

// ---- adjoint
template <int ND, class R = double>
struct Adj1 {
    R a[ND];
    R p;
};
template <int ND, class R = double>
SSDE_HD Adj1<ND, R> adj_zero1() {
    Adj1<ND, R> r;
    r.p = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) r.a[d] = 0.0;
    return r;
}

template <int ND, class R = double>
struct BwdElem1 {
    R L;
    R z[ND];
    R D;
    static constexpr int NDBL = (2 + ND) * ScalarOf<R>::NDBL;
};
template <int ND, class R = double>
SSDE_HD BwdElem1<ND, R> bwd_identity1() {
    BwdElem1<ND, R> E;
    E.L = 1.0; E.D = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) E.z[d] = 0.0;
    return E;
}
template <int ND, class R>
SSDE_HD BwdElem1<ND, R> bwd_const1(const Adj1<ND, R>& g) {
    BwdElem1<ND, R> E;
    E.L = 0.0; E.D = g.p;
#pragma unroll
    for (int d = 0; d < ND; ++d) E.z[d] = -g.a[d];
    return E;
}
template <int ND, class R>
SSDE_HD BwdElem1<ND, R> bwd_row_elem1(const Step1<R>& sp, const StepAux1<ND, R>& ax, bool has_obs, bool cut) {
    BwdElem1<ND, R> E;
    R sw2 = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        E.z[d] = has_obs ? ax.w[d] : R(0.0);
        sw2 += ax.w[d] * ax.w[d];
    }
    E.D = has_obs ? R(0.5 * ((double)ND * ax.iF - sw2)) : R(0.0);
    E.L = cut ? R(0.0) : R(sp.t * (1.0 - ax.g));
    return E;
}
// E1 earlier rows, E2 later rows
template <int ND, class R>
SSDE_HD BwdElem1<ND, R> bwd_combine1(const BwdElem1<ND, R>& E1, const BwdElem1<ND, R>& E2) {
    BwdElem1<ND, R> Ro;
    Ro.L = E2.L * E1.L;
    Ro.D = E1.L * E1.L * E2.D + E1.D;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const R t1 = E1.L * E2.z[d];
        Ro.z[d] = t1 + E1.z[d];
        Ro.D -= t1 * E1.z[d];
    }
    return Ro;
}
template <int ND, class R>
SSDE_HD Adj1<ND, R> bwd_apply1(const BwdElem1<ND, R>& E, const Adj1<ND, R>& g) {
    Adj1<ND, R> r;
    r.p = E.L * E.L * g.p + E.D;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const R t1 = E.L * g.a[d];
        r.a[d] = t1 - E.z[d];
        r.p += t1 * E.z[d];
    }
    return r;
}

// Adjoint of the step quantities of one row given the adjoint g of the state it predicts:
//   cbar_d = g.a_d,  qbar = g.p,  tbar = sum_d g.a_d af_d + 2 g.p t pf,
// and the row's contribution to d nllk / d h (update part).
template <int ND, class R>
SSDE_HD void step_adjoint1(const Adj1<ND, R>& g, const Step1<R>& sp, const StepAux1<ND, R>& ax, const R* mu,
                           bool has_obs, R& tbar, R& qbar, R& cmbar, R* mubar, R& g_h) {
    tbar = 2.0 * g.p * sp.t * ax.pf;
    qbar = g.p;
    cmbar = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        tbar += g.a[d] * ax.af[d];
        cmbar += g.a[d] * mu[d];
        mubar[d] = sp.cm * g.a[d];
    }
    g_h = 0.0;
    if (has_obs) {
        // hbar = g^2 pf_bar - sum_d w_d g af_bar_d + (d / F - sum w^2) / 2,  af_bar = t abar+, pf_bar = t^2 pbar+
        R sw2 = 0.0, acc = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            acc += ax.w[d] * ax.g * (sp.t * g.a[d]);
            sw2 += ax.w[d] * ax.w[d];
        }
        g_h = ax.g * ax.g * (sp.t * sp.t * g.p) - acc + 0.5 * ((double)ND * ax.iF - sw2);
    }
}

}  // namespace ssde
