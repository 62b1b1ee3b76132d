// TEST INFRASTRUCTURE ONLY.  Host-compiled harness around smoothsde_b200/csrc/ctcrw_math.cuh so
// that the scan algebra (append / combine / apply, forward and adjoint) can be checked against
// the oracle on a machine without a GPU.  It emulates the kernels' decomposition (rows -> per
// thread chunks of `lc` rows -> tiles of `nt` threads -> chained tile prefixes) sequentially.
// Nothing in the product path links or loads this file.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../smoothsde_b200/csrc/ctcrw_math.cuh"

using namespace ssde;

enum { FLAG_START = 1, FLAG_LAST = 2, FLAG_OBS = 4 };

template <class R> static R mk(double v, double d);
template <> double mk<double>(double v, double) { return v; }
template <> Dual mk<Dual>(double v, double d) { return Dual(v, d); }

template <int ND, class R>
struct Row {
    StepParT<R> sp;
    double y[ND], dt;
    R mu[ND], tau, e, s2;
    bool start, last, obs;
    int track;
};

// eta_dot: direction in linear-predictor space (same shape as eta), nullptr for R = double
template <int ND, class R>
static std::vector<Row<ND, R>> build_rows(int64_t n, const uint8_t* flags, const double* y,
                                          const double* dt, const double* eta, const double* eta_dot) {
    std::vector<Row<ND, R>> rows(n);
    int track = -1;
    for (int64_t i = 0; i < n; ++i) {
        Row<ND, R>& r = rows[i];
        auto E = [&](int c) { return mk<R>(eta[i * (ND + 2) + c], eta_dot ? eta_dot[i * (ND + 2) + c] : 0.0); };
        r.start = flags[i] & FLAG_START;
        r.last = flags[i] & FLAG_LAST;
        r.obs = flags[i] & FLAG_OBS;
        if (r.start) ++track;
        r.track = track;
        for (int d = 0; d < ND; ++d) { r.y[d] = y[i * ND + d]; r.mu[d] = E(d); }
        r.dt = dt[i];
        transform_row(E(ND), E(ND + 1), r.dt, r.tau, r.e, r.s2);
        r.sp = make_step(r.tau, r.e, r.s2, r.dt);
    }
    return rows;
}

template <int ND, class R>
static State<ND, R> start_state(const double* a0, const double* P0, int track) {
    State<ND, R> s;
    for (int d = 0; d < ND; ++d) s.a[d] = {a0[track * 2 * ND + 2 * d], a0[track * 2 * ND + 2 * d + 1]};
    s.P = {P0[0], P0[1], P0[2]};
    return s;
}

// mode 0: plain sequential filter + sequential adjoint.
// mode 1: emulated chunked scan (lc rows per thread, nt threads per tile).
template <int ND, class R>
static int run(int mode, int64_t n, const uint8_t* flags, const double* y, const double* dt,
               const double* eta, const double* eta_dot, const double* a0, const double* P0, R h, int lc, int nt,
               double* out_llk, double* eta_bar, double* eta_bar_dot, double* out_gh, double* aest) {
    auto rows = build_rows<ND, R>(n, flags, y, dt, eta, eta_dot);
    std::vector<State<ND, R>> pre(n);       // predicted state of each row (state BEFORE the row)
    R llk = 0.0;
    if (mode == 0) {
        State<ND, R> s = start_state<ND, R>(a0, P0, 0);
        for (int64_t i = 0; i < n; ++i) {
            pre[i] = s;
            if (rows[i].start) { s = start_state<ND, R>(a0, P0, rows[i].track); continue; }
            llk += fwd_step<ND, false>(s, rows[i].sp, rows[i].y, rows[i].mu, rows[i].obs, h, nullptr);
        }
    } else {
        const int64_t chunk = lc, tile = (int64_t)lc * nt;
        const int64_t nchunks = (n + chunk - 1) / chunk;
        std::vector<FwdElem<ND, R>> agg(nchunks);
        for (int64_t c = 0; c < nchunks; ++c) {
            FwdElem<ND, R> E = fwd_identity<ND, R>();
            for (int64_t i = c * chunk; i < (c + 1) * chunk && i < n; ++i) {
                if (rows[i].start) fwd_append_start<ND>(E, start_state<ND, R>(a0, P0, rows[i].track));
                else fwd_append<ND>(E, rows[i].sp, rows[i].y, rows[i].mu, rows[i].obs, h);
            }
            agg[c] = E;
        }
        // tile-level: exclusive prefix of chunk aggregates inside a tile via fwd_combine, tile
        // prefixes chained across tiles (what the look-back computes)
        FwdElem<ND, R> tile_prefix = fwd_identity<ND, R>();
        State<ND, R> s_in = start_state<ND, R>(a0, P0, 0);   // irrelevant: row 0 is a start row
        for (int64_t t0 = 0; t0 < n; t0 += tile) {
            FwdElem<ND, R> run_ = fwd_identity<ND, R>();
            const State<ND, R> s_tile = fwd_apply<ND>(tile_prefix, s_in);
            for (int64_t c = t0 / chunk; c < (t0 + tile) / chunk && c < nchunks; ++c) {
                State<ND, R> s = fwd_apply<ND>(run_, s_tile);
                for (int64_t i = c * chunk; i < (c + 1) * chunk && i < n; ++i) {
                    pre[i] = s;
                    if (rows[i].start) { s = start_state<ND, R>(a0, P0, rows[i].track); continue; }
                    llk += fwd_step<ND, false>(s, rows[i].sp, rows[i].y, rows[i].mu, rows[i].obs, h,
                                               nullptr);
                }
                run_ = fwd_combine<ND>(run_, agg[c]);
            }
            tile_prefix = fwd_combine<ND>(tile_prefix, run_);
        }
    }
    out_llk[0] = value(llk); out_llk[1] = tangent(llk);
    if (aest) {
        // REPORT(aest_all): row i holds the state AFTER iteration i (nllk_ctcrw.hpp:246)
        for (int64_t i = 0; i < n; ++i) {
            State<ND, R> s = pre[i];
            if (rows[i].start) s = start_state<ND, R>(a0, P0, rows[i].track);
            else fwd_step<ND, false>(s, rows[i].sp, rows[i].y, rows[i].mu, rows[i].obs, h, nullptr);
            for (int d = 0; d < ND; ++d) { aest[i * 2 * ND + 2 * d] = value(s.a[d].x); aest[i * 2 * ND + 2 * d + 1] = value(s.a[d].y); }
        }
    }
    if (!eta_bar) return 0;

    // ---- adjoint ----
    R gh = 0.0;
    std::memset(eta_bar, 0, sizeof(double) * n * (ND + 2));
    if (eta_bar_dot) std::memset(eta_bar_dot, 0, sizeof(double) * n * (ND + 2));
    auto put = [&](int64_t i, int c, const R& g) {
        eta_bar[i * (ND + 2) + c] = value(g);
        if (eta_bar_dot) eta_bar_dot[i * (ND + 2) + c] = tangent(g);
    };
    auto row_back = [&](int64_t i, Adj<ND, R>& g) {
        const Row<ND, R>& r = rows[i];
        if (r.start) { g = adj_zero<ND, R>(); return; }
        State<ND, R> s = pre[i];
        StepAux<ND, R> ax;
        fwd_step<ND, true>(s, r.sp, r.y, r.mu, r.obs, h, &ax);
        const Adj<ND, R> gin = r.last ? adj_zero<ND, R>() : g;
        R gmu[ND], gt, gn, g_h;
        row_param_grad<ND>(gin, r.sp, ax, r.mu, r.tau, r.e, r.s2, r.dt, r.obs, gmu, gt, gn, g_h);
        for (int d = 0; d < ND; ++d) put(i, d, gmu[d]);
        put(i, ND, gt);
        put(i, ND + 1, gn);
        gh += g_h;
        g = bwd_apply<ND>(bwd_row_elem<ND>(r.sp, ax, r.obs, r.last), g);
    };
    if (mode == 0) {
        Adj<ND, R> g = adj_zero<ND, R>();
        for (int64_t i = n - 1; i >= 0; --i) row_back(i, g);
    } else {
        const int64_t chunk = lc, tile = (int64_t)lc * nt;
        const int64_t nchunks = (n + chunk - 1) / chunk;
        std::vector<BwdElem<ND, R>> agg(nchunks);
        for (int64_t c = 0; c < nchunks; ++c) {
            BwdElem<ND, R> E = bwd_identity<ND, R>();
            for (int64_t i = c * chunk; i < (c + 1) * chunk && i < n; ++i) {
                const Row<ND, R>& r = rows[i];
                BwdElem<ND, R> el;
                if (r.start) el = bwd_const<ND>(adj_zero<ND, R>());
                else {
                    State<ND, R> s = pre[i];
                    StepAux<ND, R> ax;
                    fwd_step<ND, true>(s, r.sp, r.y, r.mu, r.obs, h, &ax);
                    el = bwd_row_elem<ND>(r.sp, ax, r.obs, r.last);
                }
                E = bwd_combine<ND>(E, el);
            }
            agg[c] = E;
        }
        BwdElem<ND, R> suffix = bwd_identity<ND, R>();    // composite of all rows after the current tile
        const Adj<ND, R> g_end = adj_zero<ND, R>();
        const int64_t ntiles = (n + tile - 1) / tile;
        for (int64_t t = ntiles - 1; t >= 0; --t) {
            const int64_t t0 = t * tile;
            BwdElem<ND, R> run_ = bwd_identity<ND, R>();  // composite of later chunks inside this tile
            const Adj<ND, R> g_tile = bwd_apply<ND>(suffix, g_end);
            int64_t c_hi = (t0 + tile) / chunk; if (c_hi > nchunks) c_hi = nchunks;
            for (int64_t c = c_hi - 1; c >= t0 / chunk; --c) {
                Adj<ND, R> g = bwd_apply<ND>(run_, g_tile);
                int64_t i_hi = (c + 1) * chunk; if (i_hi > n) i_hi = n;
                for (int64_t i = i_hi - 1; i >= c * chunk; --i) row_back(i, g);
                run_ = bwd_combine<ND>(agg[c], run_);
            }
            suffix = bwd_combine<ND>(run_, suffix);
        }
    }
    out_gh[0] = value(gh); out_gh[1] = tangent(gh);
    return 0;
}

// out_llk[2] = (llk, d llk); out_gh[2] = (d nllk / d h, its tangent).  eta_dot == nullptr: plain doubles.
extern "C" int harness_ctcrw(int nd, int mode, int64_t n, const uint8_t* flags, const double* y,
                             const double* dt, const double* eta, const double* a0,
                             const double* P0, double h, int lc, int nt, double* out_llk,
                             double* eta_bar, double* out_gh, double* aest) {
    double llk2[2] = {0, 0}, gh2[2] = {0, 0};
    int rc = 1;
    switch (nd) {
        case 1: rc = run<1, double>(mode, n, flags, y, dt, eta, nullptr, a0, P0, h, lc, nt, llk2, eta_bar, nullptr, gh2, aest); break;
        case 2: rc = run<2, double>(mode, n, flags, y, dt, eta, nullptr, a0, P0, h, lc, nt, llk2, eta_bar, nullptr, gh2, aest); break;
        case 3: rc = run<3, double>(mode, n, flags, y, dt, eta, nullptr, a0, P0, h, lc, nt, llk2, eta_bar, nullptr, gh2, aest); break;
    }
    *out_llk = llk2[0];
    if (out_gh) *out_gh = gh2[0];
    return rc;
}

// Tangent (Dual) run of the same algebra: direction (eta_dot, h_dot).
extern "C" int harness_ctcrw_tangent(int nd, int mode, int64_t n, const uint8_t* flags, const double* y,
                                     const double* dt, const double* eta, const double* eta_dot,
                                     const double* a0, const double* P0, double h, double h_dot, int lc,
                                     int nt, double* out_llk2, double* eta_bar, double* eta_bar_dot,
                                     double* out_gh2) {
    const Dual hh(h, h_dot);
    switch (nd) {
        case 1: return run<1, Dual>(mode, n, flags, y, dt, eta, eta_dot, a0, P0, hh, lc, nt, out_llk2, eta_bar, eta_bar_dot, out_gh2, nullptr);
        case 2: return run<2, Dual>(mode, n, flags, y, dt, eta, eta_dot, a0, P0, hh, lc, nt, out_llk2, eta_bar, eta_bar_dot, out_gh2, nullptr);
    }
    return 1;
}
