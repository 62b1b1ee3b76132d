// Model traits: everything the scan kernels (kernels_ctcrw.cuh) need to know about a Kalman
// model -- the reference's three filter templates share one loop skeleton
// (nllk_ctcrw.hpp:195-247, nllk_ou_ssm.hpp:163-213, nllk_bm_ssm.hpp:127-175) and differ only in
// the state dimension and in makeT / makeQ / makeB:
//   CtcrwModel  2 states per dimension (position, velocity)   ctcrw_math.cuh
//   OuSsmModel  1 state per dimension, T = exp(-dt/tau)       ssm1_math.cuh
//   BmSsmModel  1 state per dimension, T = 1, drift mu dt     ssm1_math.cuh
// A traits class M<ND, R> provides the element types and algebra, the natural-scale transform of
// a linear-predictor row, packing of states / step quantities into arrays (accessor `at(i)`
// returns a reference to the i-th scalar), and the closed-form gradient of one row.
#pragma once

#include "ctcrw_math.cuh"
#include "dense_math.cuh"
#include "ssm1_math.cuh"

namespace ssde {

// Prior covariance of a track's first filtered state (P0, R/sde.R:551-556,582-588): `blk` is the
// shared block of the decoupled models (CTCRW: 2x2 (p11, p12, p22); SSM: p11 = c of P0 = c I),
// `dense` the packed upper triangle of the full N x N matrix for the coupled models.
struct PriorCov {
    Sym2 blk;
    double dense[10];
};

// Kernel geometry of the decoupled models (threads per CTA, CTAs per SM the kernels are compiled for)
#ifndef SSDE_KNT
#define SSDE_KNT 128
#endif
#ifndef SSDE_MINB
#define SSDE_MINB 3
#endif
// resident CTAs per SM of the forward kernel alone (it needs little shared memory once the step cache is gone)
#ifndef SSDE_FWD_MINB
#define SSDE_FWD_MINB 4
#endif
constexpr int KNT_DEFAULT = SSDE_KNT, MINB_DEFAULT = SSDE_MINB, FMINB_DEFAULT = SSDE_FWD_MINB;

constexpr double CONST_MAP_TOL = 1e-60;     // see FwdOps / is_const in common.cuh
// The device reads the threshold from constant memory so that tests can switch the constant-map
// shortcut off (ssde_debug_const_map_tol(device, -1)) and compare both paths on the same data.
#ifdef __CUDACC__
__constant__ double c_const_map_tol = CONST_MAP_TOL;
#endif
SSDE_HD double const_map_tol() {
#ifdef __CUDA_ARCH__
    return c_const_map_tol;
#else
    return CONST_MAP_TOL;
#endif
}
SSDE_HD bool tiny(double x) { return fabs(x) <= const_map_tol(); }
SSDE_HD bool tiny(const Dual& x) { const double t = const_map_tol(); return fabs(x.v) <= t && fabs(x.d) <= t; }
template <class R>
SSDE_HD bool tiny(const Mat2T<R>& m) { return tiny(m.m11) && tiny(m.m12) && tiny(m.m21) && tiny(m.m22); }

// ---------------------------------------------------------------------------------------------
template <int ND_, class R_>
struct CtcrwModel {
    static constexpr int ND = ND_;
    using R = R_;
    static constexpr int NP = ND + 2;        // SDE parameters per row: mu_1..mu_d, tau, nu
    static constexpr int SD = 2 * ND;        // state means per row (columns of a0 / aest_all)
    static constexpr int FS = 2 * ND + 3;    // scalars of a State / Adj
    static constexpr int NW = 3;             // transformed parameters kept for the adjoint kernel
    static constexpr int SPD = 2;            // states per dimension
    static constexpr int KNT = KNT_DEFAULT, MINB = MINB_DEFAULT, FMINB = FMINB_DEFAULT;
    static constexpr int LOGF_MULT = ND;     // log|F| = n_dim log F
    using Hc = R;                            // measurement covariance of a row: H = h I
    static SSDE_HD Hc row_h(const R& h, const double*, size_t, int64_t) { return h; }
    using State = ssde::State<ND, R>;
    using Adj = ssde::Adj<ND, R>;
    using FwdElem = ssde::FwdElem<ND, R>;
    using BwdElem = ssde::BwdElem<ND, R>;
    using Step = StepParT<R>;
    using Aux = StepAux<ND, R>;
    struct RowPar { R tau, e, s2; };

    static SSDE_HD RowPar transform(const R* eta, double dt) {
        RowPar r;
        transform_row(eta[ND], eta[ND + 1], dt, r.tau, r.e, r.s2);
        return r;
    }
    template <class F> static SSDE_HD void store_rowpar(const RowPar& r, F at) { at(0) = r.tau; at(1) = r.e; at(2) = r.s2; }
    template <class F> static SSDE_HD RowPar load_rowpar(F at) { return RowPar{at(0), at(1), at(2)}; }
    static SSDE_HD Step make_step(const RowPar& r, double dt) { return ssde::make_step(r.tau, r.e, r.s2, dt); }
    // a0 row = (x, 0, y, 0, ...) R/sde.R:574-580; P0 = shared 2x2 block
    static SSDE_HD State start_state(const double* a0row, const PriorCov& P0) {
        State s;
#pragma unroll
        for (int d = 0; d < ND; ++d) s.a[d] = {a0row[2 * d], a0row[2 * d + 1]};
        s.P = {P0.blk.a, P0.blk.b, P0.blk.c};
        return s;
    }
    static SSDE_HD State zero_state(const PriorCov& P0) {
        State s;
#pragma unroll
        for (int d = 0; d < ND; ++d) s.a[d] = {0.0, 0.0};
        s.P = {P0.blk.a, P0.blk.b, P0.blk.c};
        return s;
    }
    template <class F> static SSDE_HD void store_state(const State& s, F at) {
#pragma unroll
        for (int d = 0; d < ND; ++d) { at(2 * d) = s.a[d].x; at(2 * d + 1) = s.a[d].y; }
        at(2 * ND) = s.P.a; at(2 * ND + 1) = s.P.b; at(2 * ND + 2) = s.P.c;
    }
    template <class F> static SSDE_HD State load_state(F at) {
        State s;
#pragma unroll
        for (int d = 0; d < ND; ++d) s.a[d] = {at(2 * d), at(2 * d + 1)};
        s.P = {at(2 * ND), at(2 * ND + 1), at(2 * ND + 2)};
        return s;
    }
    template <class F> static SSDE_HD void store_adj(const Adj& g, F at) {
#pragma unroll
        for (int d = 0; d < ND; ++d) { at(2 * d) = g.a[d].x; at(2 * d + 1) = g.a[d].y; }
        at(2 * ND) = g.P.a; at(2 * ND + 1) = g.P.b; at(2 * ND + 2) = g.P.c;
    }
    template <class F> static SSDE_HD Adj load_adj(F at) {
        Adj g;
#pragma unroll
        for (int d = 0; d < ND; ++d) g.a[d] = {at(2 * d), at(2 * d + 1)};
        g.P = {at(2 * ND), at(2 * ND + 1), at(2 * ND + 2)};
        return g;
    }
    static SSDE_HD void store_mean(const State& s, double* o) {
#pragma unroll
        for (int d = 0; d < ND; ++d) { o[2 * d] = value(s.a[d].x); o[2 * d + 1] = value(s.a[d].y); }
    }
    // forward
    static SSDE_HD FwdElem fwd_identity() { return ssde::fwd_identity<ND, R>(); }
    static SSDE_HD void fwd_append(FwdElem& E, const Step& sp, const double* y, const R* mu, bool has, const R& h) { ssde::fwd_append<ND>(E, sp, y, mu, has, h); }
    static SSDE_HD void fwd_append_start(FwdElem& E, const State& s0) { ssde::fwd_append_start<ND>(E, s0); }
    static SSDE_HD FwdElem fwd_combine(const FwdElem& a, const FwdElem& b) { return ssde::fwd_combine<ND>(a, b); }
    static SSDE_HD State fwd_apply(const FwdElem& E, const State& s) { return ssde::fwd_apply<ND>(E, s); }
    static SSDE_HD bool fwd_is_const(const FwdElem& E) { return tiny(E.A); }
    template <bool AUX>
    static SSDE_HD void fwd_step(State& s, const Step& sp, const double* y, const R* mu, bool has, const R& h, Aux* aux, R& F, R& qd) {
        ssde::fwd_step_q<ND, AUX>(s, sp, y, mu, has, h, aux, F, qd);
    }
    // adjoint
    static SSDE_HD Adj adj_zero() { return ssde::adj_zero<ND, R>(); }
    static SSDE_HD BwdElem bwd_identity() { return ssde::bwd_identity<ND, R>(); }
    static SSDE_HD BwdElem bwd_const(const Adj& g) { return ssde::bwd_const<ND>(g); }
    static SSDE_HD BwdElem bwd_row_elem(const Step& sp, const Aux& ax, bool has, bool cut) { return ssde::bwd_row_elem<ND>(sp, ax, has, cut); }
    static SSDE_HD BwdElem bwd_combine(const BwdElem& a, const BwdElem& b) { return ssde::bwd_combine<ND>(a, b); }
    static SSDE_HD Adj bwd_apply(const BwdElem& E, const Adj& g) { return ssde::bwd_apply<ND>(E, g); }
    // one row appended to / applied from its step quantities: the row element's structural zeros are skipped
    static SSDE_HD BwdElem bwd_append_row(const BwdElem& E, const Step& sp, const Aux& ax, bool has, bool cut) {
        return ssde::bwd_append_row<ND>(E, sp, ax, has, cut);
    }
    static SSDE_HD Adj bwd_apply_row(const Step& sp, const Aux& ax, bool has, bool cut, const Adj& g) {
        return ssde::bwd_apply_row<ND>(sp, ax, has, cut, g);
    }
    static SSDE_HD bool bwd_is_const(const BwdElem& E) { return tiny(E.L); }
    // gp[NP] = d nllk / d eta of this row
    static SSDE_HD void row_param_grad(const Adj& g, const Step& sp, const Aux& ax, const R* mu, const RowPar& rp, double dt,
                                       bool has, const Hc&, R* gp, R& g_h) {
        ssde::row_param_grad<ND>(g, sp, ax, mu, rp.tau, rp.e, rp.s2, dt, has, gp, gp[ND], gp[ND + 1], g_h);
    }
    // block form of a step for the coupled filter (dense_math.cuh) and the chain rule from block
    // adjoints back to the linear predictors (gp[ND], gp[ND + 1])
    static SSDE_HD StepBlk<2, R> to_blk(const Step& sp) {
        StepBlk<2, R> k;
        k.t[0][0] = 1.0; k.t[0][1] = sp.T12; k.t[1][0] = 0.0; k.t[1][1] = sp.e;       // makeT_ctcrw :45-55
        k.q[0][0] = sp.Q.a; k.q[0][1] = sp.Q.b; k.q[1][0] = sp.Q.b; k.q[1][1] = sp.Q.c;
        k.b[0] = sp.B1; k.b[1] = sp.B2;
        return k;
    }
    static SSDE_HD void chain_blk(const StepBlk<2, R>& bar, const Step&, const RowPar& rp, double dt, R* gp) {
        ctcrw_chain(bar.t[0][1], bar.t[1][1], bar.b[0], bar.b[1], bar.q[0][0], bar.q[0][1] + bar.q[1][0], bar.q[1][1],
                    rp.tau, rp.e, rp.s2, dt, gp[ND], gp[ND + 1]);
    }
};

// ---------------------------------------------------------------------------------------------
// shared part of the one-state-per-dimension models
template <int ND_, class R_>
struct Ssm1Base {
    static constexpr int ND = ND_;
    using R = R_;
    static constexpr int SD = ND;
    static constexpr int FS = ND + 1;
    static constexpr int SPD = 1;
    static constexpr int KNT = KNT_DEFAULT, MINB = MINB_DEFAULT, FMINB = FMINB_DEFAULT;
    static constexpr int LOGF_MULT = ND;
    using Hc = R;
    static SSDE_HD Hc row_h(const R& h, const double*, size_t, int64_t) { return h; }
    using State = State1<ND, R>;
    using Adj = Adj1<ND, R>;
    using FwdElem = FwdElem1<ND, R>;
    using BwdElem = BwdElem1<ND, R>;
    using Step = Step1<R>;
    using Aux = StepAux1<ND, R>;

    // a0 row = first observation (R/sde.R:549-550); P0 = c I
    static SSDE_HD State start_state(const double* a0row, const PriorCov& P0) {
        State s;
#pragma unroll
        for (int d = 0; d < ND; ++d) s.a[d] = a0row[d];
        s.p = P0.blk.a;
        return s;
    }
    static SSDE_HD State zero_state(const PriorCov& P0) {
        State s;
#pragma unroll
        for (int d = 0; d < ND; ++d) s.a[d] = 0.0;
        s.p = P0.blk.a;
        return s;
    }
    template <class F> static SSDE_HD void store_state(const State& s, F at) {
#pragma unroll
        for (int d = 0; d < ND; ++d) at(d) = s.a[d];
        at(ND) = s.p;
    }
    template <class F> static SSDE_HD State load_state(F at) {
        State s;
#pragma unroll
        for (int d = 0; d < ND; ++d) s.a[d] = at(d);
        s.p = at(ND);
        return s;
    }
    template <class F> static SSDE_HD void store_adj(const Adj& g, F at) {
#pragma unroll
        for (int d = 0; d < ND; ++d) at(d) = g.a[d];
        at(ND) = g.p;
    }
    template <class F> static SSDE_HD Adj load_adj(F at) {
        Adj g;
#pragma unroll
        for (int d = 0; d < ND; ++d) g.a[d] = at(d);
        g.p = at(ND);
        return g;
    }
    static SSDE_HD void store_mean(const State& s, double* o) {
#pragma unroll
        for (int d = 0; d < ND; ++d) o[d] = value(s.a[d]);
    }
    static SSDE_HD FwdElem fwd_identity() { return fwd_identity1<ND, R>(); }
    static SSDE_HD void fwd_append(FwdElem& E, const Step& sp, const double* y, const R* mu, bool has, const R& h) { fwd_append1<ND>(E, sp, y, mu, has, h); }
    static SSDE_HD void fwd_append_start(FwdElem& E, const State& s0) { fwd_append_start1<ND>(E, s0); }
    static SSDE_HD FwdElem fwd_combine(const FwdElem& a, const FwdElem& b) { return fwd_combine1<ND>(a, b); }
    static SSDE_HD State fwd_apply(const FwdElem& E, const State& s) { return fwd_apply1<ND>(E, s); }
    static SSDE_HD bool fwd_is_const(const FwdElem& E) { return tiny(E.A); }
    template <bool AUX>
    static SSDE_HD void fwd_step(State& s, const Step& sp, const double* y, const R* mu, bool has, const R& h, Aux* aux, R& F, R& qd) {
        fwd_step1<ND, AUX>(s, sp, y, mu, has, h, aux, F, qd);
    }
    static SSDE_HD Adj adj_zero() { return adj_zero1<ND, R>(); }
    static SSDE_HD BwdElem bwd_identity() { return bwd_identity1<ND, R>(); }
    static SSDE_HD BwdElem bwd_const(const Adj& g) { return bwd_const1<ND>(g); }
    static SSDE_HD BwdElem bwd_row_elem(const Step& sp, const Aux& ax, bool has, bool cut) { return bwd_row_elem1<ND>(sp, ax, has, cut); }
    static SSDE_HD BwdElem bwd_combine(const BwdElem& a, const BwdElem& b) { return bwd_combine1<ND>(a, b); }
    static SSDE_HD Adj bwd_apply(const BwdElem& E, const Adj& g) { return bwd_apply1<ND>(E, g); }
    static SSDE_HD BwdElem bwd_append_row(const BwdElem& E, const Step& sp, const Aux& ax, bool has, bool cut) {
        return bwd_combine1<ND>(E, bwd_row_elem1<ND>(sp, ax, has, cut));
    }
    static SSDE_HD Adj bwd_apply_row(const Step& sp, const Aux& ax, bool has, bool cut, const Adj& g) {
        return bwd_apply1<ND>(bwd_row_elem1<ND>(sp, ax, has, cut), g);
    }
    static SSDE_HD bool bwd_is_const(const BwdElem& E) { return tiny(E.L); }
    static SSDE_HD StepBlk<1, R> to_blk(const Step& sp) {
        StepBlk<1, R> k;
        k.t[0][0] = sp.t; k.q[0][0] = sp.q; k.b[0] = sp.cm;
        return k;
    }
};

// OU + measurement error, nllk_ou_ssm.hpp: mu_d, tau = exp(eta_tau), kappa = exp(eta_kappa) (:122-124)
template <int ND_, class R_>
struct OuSsmModel : Ssm1Base<ND_, R_> {
    using B = Ssm1Base<ND_, R_>;
    using R = R_;
    static constexpr int ND = ND_;
    static constexpr int NP = ND + 2;
    static constexpr int NW = 3;
    struct RowPar { R tau, e, kappa; };
    static SSDE_HD RowPar transform(const R* eta, double dt) {
        RowPar r;
        r.tau = exp(eta[ND]);
        r.kappa = exp(eta[ND + 1]);
        r.e = exp(-dt / r.tau);                              // makeT_ou_ssm :35
        return r;
    }
    template <class F> static SSDE_HD void store_rowpar(const RowPar& r, F at) { at(0) = r.tau; at(1) = r.e; at(2) = r.kappa; }
    template <class F> static SSDE_HD RowPar load_rowpar(F at) { return RowPar{at(0), at(1), at(2)}; }
    static SSDE_HD typename B::Step make_step(const RowPar& r, double dt) {
        typename B::Step sp;
        sp.t = r.e;
        sp.cm = 1.0 - r.e;                                   // makeB_ou_ssm :50
        sp.q = r.kappa * (1.0 - exp(-2.0 * dt / r.tau));     // makeQ_ou_ssm :66 (its own exp, as in the reference)
        return sp;
    }
    // t = e, cm = 1 - e, q = kappa (1 - e2), e = exp(-dt/tau), e2 = exp(-2 dt/tau), tau = exp(eta_tau)
    static SSDE_HD void chain(const R& tbar, const R& qbar, const R& cmbar, const typename B::Step& sp, const RowPar& rp,
                              double dt, R* gp) {
        const R e2 = 1.0 - sp.q / rp.kappa;
        const R ebar = tbar - cmbar;
        const R r_ = dt / rp.tau;
        gp[ND] = ebar * rp.e * r_ - qbar * rp.kappa * e2 * (2.0 * r_);
        gp[ND + 1] = qbar * sp.q;
    }
    static SSDE_HD void row_param_grad(const typename B::Adj& g, const typename B::Step& sp, const typename B::Aux& ax, const R* mu,
                                       const RowPar& rp, double dt, bool has, const R&, R* gp, R& g_h) {
        R tbar, qbar, cmbar;
        step_adjoint1<ND>(g, sp, ax, mu, has, tbar, qbar, cmbar, gp, g_h);
        chain(tbar, qbar, cmbar, sp, rp, dt, gp);
    }
    static SSDE_HD void chain_blk(const StepBlk<1, R>& bar, const typename B::Step& sp, const RowPar& rp, double dt, R* gp) {
        chain(bar.t[0][0], bar.q[0][0], bar.b[0], sp, rp, dt, gp);
    }
};

// BM + measurement error, nllk_bm_ssm.hpp: mu_d, sigma = exp(eta_sigma) (:89-90)
template <int ND_, class R_>
struct BmSsmModel : Ssm1Base<ND_, R_> {
    using B = Ssm1Base<ND_, R_>;
    using R = R_;
    static constexpr int ND = ND_;
    static constexpr int NP = ND + 1;
    static constexpr int NW = 1;
    struct RowPar { R s2; };
    static SSDE_HD RowPar transform(const R* eta, double) {
        const R sigma = exp(eta[ND]);
        return RowPar{sigma * sigma};
    }
    template <class F> static SSDE_HD void store_rowpar(const RowPar& r, F at) { at(0) = r.s2; }
    template <class F> static SSDE_HD RowPar load_rowpar(F at) { return RowPar{at(0)}; }
    static SSDE_HD typename B::Step make_step(const RowPar& r, double dt) {
        typename B::Step sp;
        sp.t = 1.0;                                          // T = I, nllk_bm_ssm.hpp:100-101
        sp.cm = dt;                                          // drift = mu * dt, :140
        sp.q = r.s2 * dt;                                    // makeQ_bm_ssm :32
        return sp;
    }
    static SSDE_HD void row_param_grad(const typename B::Adj& g, const typename B::Step& sp, const typename B::Aux& ax, const R* mu,
                                       const RowPar&, double, bool has, const R&, R* gp, R& g_h) {
        R tbar, qbar, cmbar;
        step_adjoint1<ND>(g, sp, ax, mu, has, tbar, qbar, cmbar, gp, g_h);
        gp[ND] = 2.0 * qbar * sp.q;                          // q = exp(2 eta_sigma) dt
    }
    static SSDE_HD void chain_blk(const StepBlk<1, R>& bar, const typename B::Step& sp, const RowPar&, double, R* gp) {
        gp[ND] = 2.0 * bar.q[0][0] * sp.q;
    }
};

// ---------------------------------------------------------------------------------------------
// Coupled filter: the full N x N covariance recursion (dense_math.cuh) around the step
// quantities, natural-scale transform and chain rule of a decoupled model `Base`.  Used when the
// user supplies H_array (a general n_dim x n_dim measurement covariance per row,
// nllk_ctcrw.hpp:203-205) or a P0 that is not of the default shape.  Same kernels, wider elements.
template <class Base>
struct DenseModel {
    using R = typename Base::R;
    static constexpr int ND = Base::ND;
    static constexpr int NP = Base::NP;
    static constexpr int SD = Base::SD;
    static constexpr int SPD = Base::SPD;
    static constexpr int N = SD;
    static constexpr int NS = N * (N + 1) / 2;
    static constexpr int FS = N + NS;
    static constexpr int NW = Base::NW;
    static constexpr int KNT = 64, MINB = 1, FMINB = 1;     // wide elements: half-size CTAs, one per SM
    static constexpr int LOGF_MULT = 1;          // F_out is det F
    using State = DState<N, R>;
    using Adj = DAdj<N, R>;
    using FwdElem = DFwdElem<N, R>;
    using BwdElem = DBwdElem<N, R>;
    using Step = typename Base::Step;
    using Aux = DAux<ND, SPD, R>;
    using RowPar = typename Base::RowPar;
    using Hc = DObsCov<ND, R>;

    // H of the row at permuted position `pos`: planes of the packed upper triangle of H_array[, , i]
    // (data), or sigma_obs^2 I (makeH_ctcrw, nllk_ctcrw.hpp:30-38)
    static SSDE_HD Hc row_h(const R& h, const double* Hrow, size_t n_pad, int64_t pos) {
        Hc H;
        H.par = Hrow == nullptr;
#pragma unroll
        for (int d = 0; d < ND; ++d)
#pragma unroll
            for (int e = 0; e < ND; ++e) {
                if (Hrow) H.v[d][e] = Hrow[(size_t)sym_idx<ND>(d, e) * n_pad + pos];
                else H.v[d][e] = (d == e) ? h : R(0.0);
            }
        return H;
    }
    static SSDE_HD RowPar transform(const R* eta, double dt) { return Base::transform(eta, dt); }
    template <class F> static SSDE_HD void store_rowpar(const RowPar& r, F at) { Base::store_rowpar(r, at); }
    template <class F> static SSDE_HD RowPar load_rowpar(F at) { return Base::load_rowpar(at); }
    static SSDE_HD Step make_step(const RowPar& r, double dt) { return Base::make_step(r, dt); }

    static SSDE_HD State start_state(const double* a0row, const PriorCov& P0) {
        State s;
#pragma unroll
        for (int i = 0; i < N; ++i) s.a[i] = a0row[i];
#pragma unroll
        for (int i = 0; i < NS; ++i) s.P[i] = P0.dense[i];
        return s;
    }
    static SSDE_HD State zero_state(const PriorCov& P0) {
        State s;
#pragma unroll
        for (int i = 0; i < N; ++i) s.a[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) s.P[i] = P0.dense[i];
        return s;
    }
    template <class F> static SSDE_HD void store_state(const State& s, F at) {
#pragma unroll
        for (int i = 0; i < N; ++i) at(i) = s.a[i];
#pragma unroll
        for (int i = 0; i < NS; ++i) at(N + i) = s.P[i];
    }
    template <class F> static SSDE_HD State load_state(F at) {
        State s;
#pragma unroll
        for (int i = 0; i < N; ++i) s.a[i] = at(i);
#pragma unroll
        for (int i = 0; i < NS; ++i) s.P[i] = at(N + i);
        return s;
    }
    template <class F> static SSDE_HD void store_adj(const Adj& g, F at) {
#pragma unroll
        for (int i = 0; i < N; ++i) at(i) = g.a[i];
#pragma unroll
        for (int i = 0; i < NS; ++i) at(N + i) = g.P[i];
    }
    template <class F> static SSDE_HD Adj load_adj(F at) {
        Adj g;
#pragma unroll
        for (int i = 0; i < N; ++i) g.a[i] = at(i);
#pragma unroll
        for (int i = 0; i < NS; ++i) g.P[i] = at(N + i);
        return g;
    }
    static SSDE_HD void store_mean(const State& s, double* o) {
#pragma unroll
        for (int i = 0; i < N; ++i) o[i] = value(s.a[i]);
    }
    // forward
    static SSDE_HD FwdElem fwd_identity() { return dfwd_identity<N, R>(); }
    static SSDE_HD void fwd_append(FwdElem& E, const Step& sp, const double* y, const R* mu, bool has, const Hc& H) {
        dfwd_append<ND, SPD>(E, Base::to_blk(sp), y, mu, has, H);
    }
    static SSDE_HD void fwd_append_start(FwdElem& E, const State& s0) { dfwd_append_start<N>(E, s0); }
    static SSDE_HD FwdElem fwd_combine(const FwdElem& a, const FwdElem& b) { return dfwd_combine<N>(a, b); }
    static SSDE_HD State fwd_apply(const FwdElem& E, const State& s) { return dfwd_apply<N>(E, s); }
    static SSDE_HD bool fwd_is_const(const FwdElem& E) {
        bool c = true;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) c = c && tiny(E.A[i][j]);
        return c;
    }
    template <bool AUX>
    static SSDE_HD void fwd_step(State& s, const Step& sp, const double* y, const R* mu, bool has, const Hc& H, Aux* aux, R& F, R& qd) {
        dense_fwd_step<ND, SPD, AUX>(s, Base::to_blk(sp), y, mu, has, H, aux, F, qd);
    }
    // adjoint
    static SSDE_HD Adj adj_zero() { return dadj_zero<N, R>(); }
    static SSDE_HD BwdElem bwd_identity() { return dbwd_identity<N, R>(); }
    static SSDE_HD BwdElem bwd_const(const Adj& g) { return dbwd_const<N>(g); }
    static SSDE_HD BwdElem bwd_row_elem(const Step& sp, const Aux& ax, bool has, bool cut) {
        return dbwd_row_elem<ND, SPD>(Base::to_blk(sp), ax, has, cut);
    }
    static SSDE_HD BwdElem bwd_combine(const BwdElem& a, const BwdElem& b) { return dbwd_combine<N>(a, b); }
    static SSDE_HD Adj bwd_apply(const BwdElem& E, const Adj& g) { return dbwd_apply<N>(E, g); }
    static SSDE_HD BwdElem bwd_append_row(const BwdElem& E, const Step& sp, const Aux& ax, bool has, bool cut) {
        return dbwd_combine<N>(E, bwd_row_elem(sp, ax, has, cut));
    }
    static SSDE_HD Adj bwd_apply_row(const Step& sp, const Aux& ax, bool has, bool cut, const Adj& g) {
        return dbwd_apply<N>(bwd_row_elem(sp, ax, has, cut), g);
    }
    static SSDE_HD bool bwd_is_const(const BwdElem& E) {
        bool c = true;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) c = c && tiny(E.L[i][j]);
        return c;
    }
    static SSDE_HD void row_param_grad(const Adj& g, const Step& sp, const Aux& ax, const R* mu, const RowPar& rp, double dt,
                                       bool has, const Hc& H, R* gp, R& g_h) {
        StepBlk<SPD, R> bar;
        dense_step_adjoint<ND, SPD>(g, Base::to_blk(sp), ax, mu, has, H.par, bar, gp, g_h);
        Base::chain_blk(bar, sp, rp, dt, gp);
    }
};

}  // namespace ssde
