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

template <int ND>
struct Row {
    StepPar sp;
    double y[ND], mu[ND], tau, e, s2, dt;
    bool start, last, obs;
    int track;
};

template <int ND>
static std::vector<Row<ND>> build_rows(int64_t n, const uint8_t* flags, const double* y,
                                       const double* dt, const double* eta) {
    std::vector<Row<ND>> rows(n);
    int track = -1;
    for (int64_t i = 0; i < n; ++i) {
        Row<ND>& r = rows[i];
        r.start = flags[i] & FLAG_START;
        r.last = flags[i] & FLAG_LAST;
        r.obs = flags[i] & FLAG_OBS;
        if (r.start) ++track;
        r.track = track;
        for (int d = 0; d < ND; ++d) { r.y[d] = y[i * ND + d]; r.mu[d] = eta[i * (ND + 2) + d]; }
        r.dt = dt[i];
        transform_row(eta[i * (ND + 2) + ND], eta[i * (ND + 2) + ND + 1], r.dt, r.tau, r.e, r.s2);
        r.sp = make_step(r.tau, r.e, r.s2, r.dt);
    }
    return rows;
}

template <int ND>
static State<ND> start_state(const double* a0, const double* P0, int track) {
    State<ND> s;
    for (int d = 0; d < ND; ++d) s.a[d] = {a0[track * 2 * ND + 2 * d], a0[track * 2 * ND + 2 * d + 1]};
    s.P = {P0[0], P0[1], P0[2]};
    return s;
}

// mode 0: plain sequential filter + sequential adjoint.
// mode 1: emulated chunked scan (lc rows per thread, nt threads per tile).
template <int ND>
static int run(int mode, int64_t n, const uint8_t* flags, const double* y, const double* dt,
               const double* eta, const double* a0, const double* P0, double h, int lc, int nt,
               double* out_llk, double* eta_bar, double* out_gh, double* aest) {
    auto rows = build_rows<ND>(n, flags, y, dt, eta);
    std::vector<State<ND>> pre(n);       // predicted state of each row (state BEFORE the row)
    double llk = 0.0;
    if (mode == 0) {
        State<ND> s = start_state<ND>(a0, P0, 0);
        for (int64_t i = 0; i < n; ++i) {
            pre[i] = s;
            if (rows[i].start) { s = start_state<ND>(a0, P0, rows[i].track); continue; }
            llk += fwd_step<ND, false>(s, rows[i].sp, rows[i].y, rows[i].mu, rows[i].obs, h, nullptr);
        }
    } else {
        const int64_t chunk = lc, tile = (int64_t)lc * nt;
        const int64_t nchunks = (n + chunk - 1) / chunk;
        std::vector<FwdElem<ND>> agg(nchunks);
        for (int64_t c = 0; c < nchunks; ++c) {
            FwdElem<ND> E = fwd_identity<ND>();
            for (int64_t i = c * chunk; i < (c + 1) * chunk && i < n; ++i) {
                if (rows[i].start) fwd_append_start<ND>(E, start_state<ND>(a0, P0, rows[i].track));
                else fwd_append<ND>(E, rows[i].sp, rows[i].y, rows[i].mu, rows[i].obs, h);
            }
            agg[c] = E;
        }
        // tile-level: exclusive prefix of chunk aggregates inside a tile via fwd_combine, tile
        // prefixes chained across tiles (what the look-back computes)
        FwdElem<ND> tile_prefix = fwd_identity<ND>();
        State<ND> s_in = start_state<ND>(a0, P0, 0);   // irrelevant: row 0 is a start row
        for (int64_t t0 = 0; t0 < n; t0 += tile) {
            FwdElem<ND> run_ = fwd_identity<ND>();
            const State<ND> s_tile = fwd_apply<ND>(tile_prefix, s_in);
            for (int64_t c = t0 / chunk; c < (t0 + tile) / chunk && c < nchunks; ++c) {
                State<ND> s = fwd_apply<ND>(run_, s_tile);
                for (int64_t i = c * chunk; i < (c + 1) * chunk && i < n; ++i) {
                    pre[i] = s;
                    if (rows[i].start) { s = start_state<ND>(a0, P0, rows[i].track); continue; }
                    llk += fwd_step<ND, false>(s, rows[i].sp, rows[i].y, rows[i].mu, rows[i].obs, h,
                                               nullptr);
                }
                run_ = fwd_combine<ND>(run_, agg[c]);
            }
            tile_prefix = fwd_combine<ND>(tile_prefix, run_);
        }
    }
    *out_llk = llk;
    if (aest) {
        // REPORT(aest_all): row i holds the state AFTER iteration i (nllk_ctcrw.hpp:246)
        for (int64_t i = 0; i < n; ++i) {
            State<ND> s = pre[i];
            if (rows[i].start) s = start_state<ND>(a0, P0, rows[i].track);
            else fwd_step<ND, false>(s, rows[i].sp, rows[i].y, rows[i].mu, rows[i].obs, h, nullptr);
            for (int d = 0; d < ND; ++d) { aest[i * 2 * ND + 2 * d] = s.a[d].x; aest[i * 2 * ND + 2 * d + 1] = s.a[d].y; }
        }
    }
    if (!eta_bar) return 0;

    // ---- adjoint ----
    double gh = 0.0;
    std::memset(eta_bar, 0, sizeof(double) * n * (ND + 2));
    auto row_back = [&](int64_t i, Adj<ND>& g) {
        const Row<ND>& r = rows[i];
        if (r.start) { g = adj_zero<ND>(); return; }
        State<ND> s = pre[i];
        StepAux<ND> ax;
        fwd_step<ND, true>(s, r.sp, r.y, r.mu, r.obs, h, &ax);
        const Adj<ND> gin = r.last ? adj_zero<ND>() : g;
        double gmu[ND], gt, gn, g_h;
        row_param_grad<ND>(gin, r.sp, ax, r.mu, r.tau, r.e, r.s2, r.dt, r.obs, gmu, gt, gn, g_h);
        for (int d = 0; d < ND; ++d) eta_bar[i * (ND + 2) + d] = gmu[d];
        eta_bar[i * (ND + 2) + ND] = gt;
        eta_bar[i * (ND + 2) + ND + 1] = gn;
        gh += g_h;
        g = bwd_apply<ND>(bwd_row_elem<ND>(r.sp, ax, r.obs, r.last), g);
    };
    if (mode == 0) {
        Adj<ND> g = adj_zero<ND>();
        for (int64_t i = n - 1; i >= 0; --i) row_back(i, g);
    } else {
        const int64_t chunk = lc, tile = (int64_t)lc * nt;
        const int64_t nchunks = (n + chunk - 1) / chunk;
        std::vector<BwdElem<ND>> agg(nchunks);
        for (int64_t c = 0; c < nchunks; ++c) {
            BwdElem<ND> E = bwd_identity<ND>();
            for (int64_t i = c * chunk; i < (c + 1) * chunk && i < n; ++i) {
                const Row<ND>& r = rows[i];
                BwdElem<ND> el;
                if (r.start) el = bwd_const<ND>(adj_zero<ND>());
                else {
                    State<ND> s = pre[i];
                    StepAux<ND> ax;
                    fwd_step<ND, true>(s, r.sp, r.y, r.mu, r.obs, h, &ax);
                    el = bwd_row_elem<ND>(r.sp, ax, r.obs, r.last);
                }
                E = bwd_combine<ND>(E, el);
            }
            agg[c] = E;
        }
        BwdElem<ND> suffix = bwd_identity<ND>();    // composite of all rows after the current tile
        const Adj<ND> g_end = adj_zero<ND>();
        const int64_t ntiles = (n + tile - 1) / tile;
        for (int64_t t = ntiles - 1; t >= 0; --t) {
            const int64_t t0 = t * tile;
            BwdElem<ND> run_ = bwd_identity<ND>();  // composite of later chunks inside this tile
            const Adj<ND> g_tile = bwd_apply<ND>(suffix, g_end);
            int64_t c_hi = (t0 + tile) / chunk; if (c_hi > nchunks) c_hi = nchunks;
            for (int64_t c = c_hi - 1; c >= t0 / chunk; --c) {
                Adj<ND> g = bwd_apply<ND>(run_, g_tile);
                int64_t i_hi = (c + 1) * chunk; if (i_hi > n) i_hi = n;
                for (int64_t i = i_hi - 1; i >= c * chunk; --i) row_back(i, g);
                run_ = bwd_combine<ND>(agg[c], run_);
            }
            suffix = bwd_combine<ND>(run_, suffix);
        }
    }
    *out_gh = gh;
    return 0;
}

extern "C" int harness_ctcrw(int nd, int mode, int64_t n, const uint8_t* flags, const double* y,
                             const double* dt, const double* eta, const double* a0,
                             const double* P0, double h, int lc, int nt, double* out_llk,
                             double* eta_bar, double* out_gh, double* aest) {
    switch (nd) {
        case 1: return run<1>(mode, n, flags, y, dt, eta, a0, P0, h, lc, nt, out_llk, eta_bar, out_gh, aest);
        case 2: return run<2>(mode, n, flags, y, dt, eta, a0, P0, h, lc, nt, out_llk, eta_bar, out_gh, aest);
        case 3: return run<3>(mode, n, flags, y, dt, eta, a0, P0, h, lc, nt, out_llk, eta_bar, out_gh, aest);
    }
    return 1;
}

