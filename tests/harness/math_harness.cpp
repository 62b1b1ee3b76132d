// TEST INFRASTRUCTURE ONLY.  Host-compiled harness around smoothsde_b200/csrc/ctcrw_math.cuh so
// that the scan algebra (append / combine / apply, forward and adjoint) can be checked against
// the oracle on a machine without a GPU.  It emulates the kernels' decomposition (rows -> per
// thread chunks of `lc` rows -> tiles of `nt` threads -> chained tile prefixes) sequentially.
// Nothing in the product path links or loads this file.
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <vector>

#include "../../smoothsde_b200/csrc/models.cuh"

using namespace ssde;

enum { FLAG_START = 1, FLAG_LAST = 2, FLAG_OBS = 4 };

template <class R> static R mk(double v, double d);
template <> double mk<double>(double v, double) { return v; }
template <> Dual mk<Dual>(double v, double d) { return Dual(v, d); }

template <class M>
struct Row {
    typename M::Step sp;
    typename M::RowPar rp;
    double y[M::ND], dt;
    typename M::R mu[M::ND];
    bool start, last, obs;
    int track;
};

// eta_dot: direction in linear-predictor space (same shape as eta), nullptr for R = double
template <class M>
static std::vector<Row<M>> build_rows(int64_t n, const uint8_t* flags, const double* y,
                                      const double* dt, const double* eta, const double* eta_dot) {
    using R = typename M::R;
    constexpr int ND = M::ND, NP = M::NP;
    std::vector<Row<M>> rows(n);
    int track = -1;
    for (int64_t i = 0; i < n; ++i) {
        Row<M>& r = rows[i];
        auto E = [&](int c) { return mk<R>(eta[i * NP + c], eta_dot ? eta_dot[i * NP + c] : 0.0); };
        r.start = flags[i] & FLAG_START;
        r.last = flags[i] & FLAG_LAST;
        r.obs = flags[i] & FLAG_OBS;
        if (r.start) ++track;
        r.track = track;
        for (int d = 0; d < ND; ++d) { r.y[d] = y[i * ND + d]; r.mu[d] = E(d); }
        r.dt = dt[i];
        R er[NP];
        for (int c = 0; c < NP; ++c) er[c] = E(c);
        r.rp = M::transform(er, r.dt);
        r.sp = M::make_step(r.rp, r.dt);
    }
    return rows;
}

template <class M>
static typename M::State start_state(const double* a0, const PriorCov& P0, int track) {
    return M::start_state(a0 + (size_t)track * M::SD, P0);
}
static PriorCov prior_blk(const double* P0) {
    PriorCov pc{};
    pc.blk = Sym2{P0[0], P0[1], P0[2]};
    return pc;
}

// mode 0: plain sequential filter + sequential adjoint.
// mode 1: emulated chunked scan (lc rows per thread, nt threads per tile).
template <class M>
static int run(int mode, int64_t n, const uint8_t* flags, const double* y, const double* dt,
               const double* eta, const double* eta_dot, const double* a0, const PriorCov& P0, typename M::R h, int lc, int nt,
               double* out_llk, double* eta_bar, double* eta_bar_dot, double* out_gh, double* aest,
               const double* Hplanes = nullptr) {
    using R = typename M::R;
    constexpr int ND = M::ND, NP = M::NP;
    auto rows = build_rows<M>(n, flags, y, dt, eta, eta_dot);
    // measurement covariance of row i (Hplanes: [ND(ND+1)/2][n] packed upper triangles, coupled models only)
    auto hrow = [&](int64_t i) { return M::row_h(h, Hplanes, (size_t)n, i); };
    auto step_llk = [&](typename M::State& s, const Row<M>& r, typename M::Aux* ax) -> R {
        R F, quad;
        const int64_t i = &r - rows.data();
        if (ax) M::template fwd_step<true>(s, r.sp, r.y, r.mu, r.obs, hrow(i), ax, F, quad);
        else M::template fwd_step<false>(s, r.sp, r.y, r.mu, r.obs, hrow(i), nullptr, F, quad);
        return r.obs ? R(-0.5 * ((double)M::LOGF_MULT * log(F) + quad)) : R(0.0);
    };
    std::vector<typename M::State> pre(n);       // predicted state of each row (state BEFORE the row)
    R llk = 0.0;
    if (mode == 0) {
        typename M::State s = start_state<M>(a0, P0, 0);
        for (int64_t i = 0; i < n; ++i) {
            pre[i] = s;
            if (rows[i].start) { s = start_state<M>(a0, P0, rows[i].track); continue; }
            llk += step_llk(s, rows[i], nullptr);
        }
    } else {
        const int64_t chunk = lc, tile = (int64_t)lc * nt;
        const int64_t nchunks = (n + chunk - 1) / chunk;
        std::vector<typename M::FwdElem> agg(nchunks);
        for (int64_t c = 0; c < nchunks; ++c) {
            typename M::FwdElem E = M::fwd_identity();
            for (int64_t i = c * chunk; i < (c + 1) * chunk && i < n; ++i) {
                if (rows[i].start) M::fwd_append_start(E, start_state<M>(a0, P0, rows[i].track));
                else M::fwd_append(E, rows[i].sp, rows[i].y, rows[i].mu, rows[i].obs, hrow(i));
            }
            agg[c] = E;
        }
        // tile-level: exclusive prefix of chunk aggregates inside a tile via fwd_combine, tile
        // prefixes chained across tiles (what the look-back computes)
        typename M::FwdElem tile_prefix = M::fwd_identity();
        typename M::State s_in = start_state<M>(a0, P0, 0);   // irrelevant: row 0 is a start row
        for (int64_t t0 = 0; t0 < n; t0 += tile) {
            typename M::FwdElem run_ = M::fwd_identity();
            const typename M::State s_tile = M::fwd_apply(tile_prefix, s_in);
            for (int64_t c = t0 / chunk; c < (t0 + tile) / chunk && c < nchunks; ++c) {
                typename M::State s = M::fwd_apply(run_, s_tile);
                for (int64_t i = c * chunk; i < (c + 1) * chunk && i < n; ++i) {
                    pre[i] = s;
                    if (rows[i].start) { s = start_state<M>(a0, P0, rows[i].track); continue; }
                    llk += step_llk(s, rows[i], nullptr);
                }
                run_ = M::fwd_combine(run_, agg[c]);
            }
            tile_prefix = M::fwd_combine(tile_prefix, run_);
        }
    }
    out_llk[0] = value(llk); out_llk[1] = tangent(llk);
    if (aest) {
        // REPORT(aest_all): row i holds the state AFTER iteration i (nllk_ctcrw.hpp:246)
        for (int64_t i = 0; i < n; ++i) {
            typename M::State s = pre[i];
            if (rows[i].start) s = start_state<M>(a0, P0, rows[i].track);
            else step_llk(s, rows[i], nullptr);
            M::store_mean(s, aest + i * M::SD);
        }
    }
    if (!eta_bar) return 0;

    // ---- adjoint ----
    R gh = 0.0;
    std::memset(eta_bar, 0, sizeof(double) * n * NP);
    if (eta_bar_dot) std::memset(eta_bar_dot, 0, sizeof(double) * n * NP);
    auto put = [&](int64_t i, int c, const R& g) {
        eta_bar[i * NP + c] = value(g);
        if (eta_bar_dot) eta_bar_dot[i * NP + c] = tangent(g);
    };
    auto row_back = [&](int64_t i, typename M::Adj& g) {
        const Row<M>& r = rows[i];
        if (r.start) { g = M::adj_zero(); return; }
        typename M::State s = pre[i];
        typename M::Aux ax;
        step_llk(s, r, &ax);
        const typename M::Adj gin = r.last ? M::adj_zero() : g;
        R gp[NP], g_h;
        for (int c = 0; c < NP; ++c) gp[c] = 0.0;
        M::row_param_grad(gin, r.sp, ax, r.mu, r.rp, r.dt, r.obs, hrow(i), gp, g_h);
        for (int c = 0; c < NP; ++c) put(i, c, gp[c]);
        gh += g_h;
        g = M::bwd_apply_row(r.sp, ax, r.obs, r.last, g);         // as the adjoint kernel's sweep does
    };
    if (mode == 0) {
        typename M::Adj g = M::adj_zero();
        for (int64_t i = n - 1; i >= 0; --i) row_back(i, g);
    } else {
        const int64_t chunk = lc, tile = (int64_t)lc * nt;
        const int64_t nchunks = (n + chunk - 1) / chunk;
        std::vector<typename M::BwdElem> agg(nchunks);
        for (int64_t c = 0; c < nchunks; ++c) {
            typename M::BwdElem E = M::bwd_identity();
            for (int64_t i = c * chunk; i < (c + 1) * chunk && i < n; ++i) {
                const Row<M>& r = rows[i];
                if (r.start) E = M::bwd_combine(E, M::bwd_const(M::adj_zero()));
                else {
                    typename M::State s = pre[i];
                    typename M::Aux ax;
                    step_llk(s, r, &ax);
                    E = M::bwd_append_row(E, r.sp, ax, r.obs, r.last);   // as the adjoint kernel's element phase does
                }
            }
            agg[c] = E;
        }
        typename M::BwdElem suffix = M::bwd_identity();    // composite of all rows after the current tile
        const typename M::Adj g_end = M::adj_zero();
        const int64_t ntiles = (n + tile - 1) / tile;
        for (int64_t t = ntiles - 1; t >= 0; --t) {
            const int64_t t0 = t * tile;
            typename M::BwdElem run_ = M::bwd_identity();  // composite of later chunks inside this tile
            const typename M::Adj g_tile = M::bwd_apply(suffix, g_end);
            int64_t c_hi = (t0 + tile) / chunk; if (c_hi > nchunks) c_hi = nchunks;
            for (int64_t c = c_hi - 1; c >= t0 / chunk; --c) {
                typename M::Adj g = M::bwd_apply(run_, g_tile);
                int64_t i_hi = (c + 1) * chunk; if (i_hi > n) i_hi = n;
                for (int64_t i = i_hi - 1; i >= c * chunk; --i) row_back(i, g);
                run_ = M::bwd_combine(agg[c], run_);
            }
            suffix = M::bwd_combine(run_, suffix);
        }
    }
    out_gh[0] = value(gh); out_gh[1] = tangent(gh);
    return 0;
}

// model: 0 = CTCRW, 1 = OU_SSM, 2 = BM_SSM
template <class R, class Fn>
static int with_model(int model, int nd, Fn fn) {
    if (model == 0) { if (nd == 1) return fn(CtcrwModel<1, R>{}); if (nd == 2) return fn(CtcrwModel<2, R>{}); if (nd == 3) return fn(CtcrwModel<3, R>{}); }
    if (model == 1) { if (nd == 1) return fn(OuSsmModel<1, R>{}); if (nd == 2) return fn(OuSsmModel<2, R>{}); }
    if (model == 2) { if (nd == 1) return fn(BmSsmModel<1, R>{}); if (nd == 2) return fn(BmSsmModel<2, R>{}); if (nd == 3) return fn(BmSsmModel<3, R>{}); }
    return 1;
}

// out_llk[2] = (llk, d llk); out_gh[2] = (d nllk / d h, its tangent).  eta_dot == nullptr: plain doubles.
extern "C" int harness_kalman(int model, int nd, int mode, int64_t n, const uint8_t* flags, const double* y,
                              const double* dt, const double* eta, const double* a0,
                              const double* P0, double h, int lc, int nt, double* out_llk,
                              double* eta_bar, double* out_gh, double* aest) {
    double llk2[2] = {0, 0}, gh2[2] = {0, 0};
    int rc = with_model<double>(model, nd, [&](auto m) {
        return run<decltype(m)>(mode, n, flags, y, dt, eta, nullptr, a0, prior_blk(P0), h, lc, nt, llk2, eta_bar, nullptr, gh2, aest);
    });
    *out_llk = llk2[0];
    if (out_gh) *out_gh = gh2[0];
    return rc;
}
extern "C" int harness_ctcrw(int nd, int mode, int64_t n, const uint8_t* flags, const double* y,
                             const double* dt, const double* eta, const double* a0,
                             const double* P0, double h, int lc, int nt, double* out_llk,
                             double* eta_bar, double* out_gh, double* aest) {
    return harness_kalman(0, nd, mode, n, flags, y, dt, eta, a0, P0, h, lc, nt, out_llk, eta_bar, out_gh, aest);
}

// Tangent (Dual) run of the same algebra: direction (eta_dot, h_dot).
extern "C" int harness_kalman_tangent(int model, int nd, int mode, int64_t n, const uint8_t* flags, const double* y,
                                      const double* dt, const double* eta, const double* eta_dot,
                                      const double* a0, const double* P0, double h, double h_dot, int lc,
                                      int nt, double* out_llk2, double* eta_bar, double* eta_bar_dot,
                                      double* out_gh2) {
    const Dual hh(h, h_dot);
    return with_model<Dual>(model, nd, [&](auto m) {
        return run<decltype(m)>(mode, n, flags, y, dt, eta, eta_dot, a0, prior_blk(P0), hh, lc, nt, out_llk2, eta_bar, eta_bar_dot, out_gh2, nullptr);
    });
}
extern "C" int harness_ctcrw_tangent(int nd, int mode, int64_t n, const uint8_t* flags, const double* y,
                                     const double* dt, const double* eta, const double* eta_dot,
                                     const double* a0, const double* P0, double h, double h_dot, int lc,
                                     int nt, double* out_llk2, double* eta_bar, double* eta_bar_dot,
                                     double* out_gh2) {
    return harness_kalman_tangent(0, nd, mode, n, flags, y, dt, eta, eta_dot, a0, P0, h, h_dot, lc, nt, out_llk2, eta_bar,
                                  eta_bar_dot, out_gh2);
}

// Coupled filter (DenseModel, dense_math.cuh): P0 is the full SD x SD matrix (row-major, symmetric),
// Hplanes == nullptr -> H = h I, else [ND(ND+1)/2][n] packed upper triangles of the rows' H.
template <class R, class Fn>
static int with_dense_model(int model, int nd, Fn fn) {
    if (model == 0) { if (nd == 1) return fn(DenseModel<CtcrwModel<1, R>>{}); if (nd == 2) return fn(DenseModel<CtcrwModel<2, R>>{}); }
    if (model == 1) { if (nd == 1) return fn(DenseModel<OuSsmModel<1, R>>{}); if (nd == 2) return fn(DenseModel<OuSsmModel<2, R>>{}); }
    if (model == 2) { if (nd == 1) return fn(DenseModel<BmSsmModel<1, R>>{}); if (nd == 2) return fn(DenseModel<BmSsmModel<2, R>>{}); if (nd == 3) return fn(DenseModel<BmSsmModel<3, R>>{}); }
    return 1;
}
static PriorCov prior_dense(const double* P0, int m) {
    PriorCov pc{};
    for (int r = 0; r < m; ++r)
        for (int c = r; c < m; ++c) pc.dense[r * m - r * (r - 1) / 2 + (c - r)] = P0[r * m + c];
    return pc;
}
extern "C" int harness_dense(int model, int nd, int mode, int64_t n, const uint8_t* flags, const double* y,
                             const double* dt, const double* eta, const double* eta_dot, const double* a0,
                             const double* P0, int m, const double* Hplanes, double h, double h_dot, int lc, int nt,
                             double* out_llk2, double* eta_bar, double* eta_bar_dot, double* out_gh2, double* aest) {
    const PriorCov pc = prior_dense(P0, m);
    if (eta_dot) {
        const Dual hh(h, h_dot);
        return with_dense_model<Dual>(model, nd, [&](auto mm) {
            return run<decltype(mm)>(mode, n, flags, y, dt, eta, eta_dot, a0, pc, hh, lc, nt, out_llk2, eta_bar, eta_bar_dot, out_gh2, aest, Hplanes);
        });
    }
    return with_dense_model<double>(model, nd, [&](auto mm) {
        return run<decltype(mm)>(mode, n, flags, y, dt, eta, nullptr, a0, pc, h, lc, nt, out_llk2, eta_bar, nullptr, out_gh2, aest, Hplanes);
    });
}

// dense_math.cuh helpers on their own: pivoted inverse of a general N x N matrix (row-major in / out)
extern "C" int harness_inv_general(int n, const double* X, double* Xi) {
    auto run_n = [&](auto tag) {
        constexpr int N = decltype(tag)::value;
        double a[N][N], b[N][N];
        for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) a[i][j] = X[i * N + j];
        inv_general<N>(a, b);
        for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) Xi[i * N + j] = b[i][j];
        return 0;
    };
    if (n == 1) return run_n(std::integral_constant<int, 1>{});
    if (n == 2) return run_n(std::integral_constant<int, 2>{});
    if (n == 3) return run_n(std::integral_constant<int, 3>{});
    if (n == 4) return run_n(std::integral_constant<int, 4>{});
    return 1;
}
