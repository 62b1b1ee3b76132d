// CTCRW Kalman filter as a time-parallel scan, fused with the spline linear predictor and (in
// the adjoint kernel) with the transposed design product.  Replaces the sequential loop
// nllk_ctcrw.hpp:195-247 (+ :143-156 in front of it) and TMB's reverse sweep of both.
//
// Rows of all tracks are stacked exactly as in the reference's data list; a track start is a
// "constant map" element, so the whole stack is ONE segmented scan and the same two kernels serve
// many short tracks, few long tracks and a single 1e8-row track.  Work decomposition (layout in
// design.cuh): tile = NT/32 warp-tiles, tiles taken from a dynamic ticket,
//   (1) each thread walks its LC rows: TMA-staged design values -> eta (one pass against a dense
//       per-warp coefficient table) -> natural-scale parameters (tau, e, s2: to HBM for the adjoint
//       kernel) -> step matrices in registers -> composes the rows into one scan element
//       (fwd_append),
//   (2) warp shuffle scan + cross-warp scan of the thread elements,
//   (3) chained look-back across tiles gives the state at the tile start,
//   (4) one checkpoint per thread chunk = its exact start state; only when no adjoint pass
//       follows (value-only evaluation, REPORT) each thread re-runs the plain filter over its rows
//       for the likelihood terms.
// The adjoint kernel walks the tiles in reverse with elements (L, z, D), recomputes the forward
// states of its rows from the checkpoints -- which yields the likelihood terms of a gradient
// evaluation as a by-product (llk_bwd) -- and ends with X' eta_bar.
#pragma once

#include "design.cuh"
#include "models.cuh"

// L2 prefetch hints (cp.async.bulk.prefetch.L2), measured on the B200 at 1024 x 1e5 rows
// (scripts/gpu_tune.sh): forward 2 = the design values of row-steps k+2, k+3 are requested while
// the TMA copy of row-step k+1 is issued, so that copy is an L2 hit (6.42 -> 6.23 ms); 1 = the
// whole warp-tile block at once (6.28 ms).  Adjoint 1 = the warp-tile's design block is requested
// when the reverse sweep starts, ~12 k cycles before the transposed product streams it
// (5.93 -> 5.63 ms); 2 = at the start of the tile: too early, L2 thrashes (6.69 ms).
#ifndef SSDE_BWD_PREFETCH
#define SSDE_BWD_PREFETCH 1
#endif
#ifndef SSDE_FWD_PREFETCH
#define SSDE_FWD_PREFETCH 2
#endif
// 1: at the start of a tile every warp asks for its 2 KB of each per-row plane (dt, obs, and in
// the adjoint kernel tau / e / s2) in L2
#ifndef SSDE_PLANE_PREFETCH
#define SSDE_PLANE_PREFETCH 1
#endif
// 1: the forward kernel forms a row's predictors in one branch-free pass over the staged values
// against a dense per-warp coefficient table (design.cuh, fill_theta_matrix / row_eta_dense)
#ifndef SSDE_ETA_DENSE
#define SSDE_ETA_DENSE 1
#endif
// 1: a warp keeps its theta cache when the next warp-tile uses the same column list
#ifndef SSDE_THETA_REUSE
#define SSDE_THETA_REUSE 1
#endif

namespace ssde {

// M = model traits (models.cuh): CtcrwModel / OuSsmModel / BmSsmModel <n_dim, scalar type>
template <class R = double>
struct KalmanArgs {
    DesignV2 X;
    Theta theta;               // [coeff_fe | coeff_re] (+ direction for the tangent pass)
    const double* obs;         // ND planes of n_pad doubles, permuted (design.cuh), NA replaced by 0
    const double* dt;          // [n_pad] permuted
    const uint8_t* flags;      // [n_pad] permuted, 0xff beyond the end
    const int64_t* track_starts;   // [n_tracks] sorted rows flagged ROW_START
    const double* a0;          // [n_tracks, M::SD]
    int n_tracks;
    PriorCov P0;
    const double* Hrow;        // coupled models with a user H_array: [ND(ND+1)/2, n_pad] permuted planes of the
                               // packed upper triangle of H_array[, , i] (else nullptr: H = sigma_obs^2 I)
    const double* par;         // device parameter vector; par[0] = log_sigma_obs
    const double* par_dot;     // R = Dual: direction in the parameter vector (else nullptr)
    const R* s_in;             // optional incoming state (M::FS scalars) for a continued shard
    const R* g_in;             // optional incoming adjoint (M::FS scalars)
    const int* mu_zero;        // device flag: every mu_d predictor is exactly 0 at these parameters
    R* wg;                     // [n_pad / 32][M::NW][32]: transformed parameters of every row (CTCRW: tau,
                               // e = exp(-dt/tau), s2), row-step-major, forward -> adjoint
    R* ckpt;                   // [M::FS, nchunks] start state of every thread chunk
    int64_t nchunks;           // n_pad / LC
    double* tile_llk;          // [n_pad / WT] one partial log-likelihood per warp-tile
    double* tile_gh;           // [n_pad / WT] one partial d nllk / d h per warp-tile; R = Dual: the
                               // tangents follow at tile_gh[n_pad / WT + q]
    double* grad_theta;        // [p_theta] (R = Dual: [2 p_theta], tangents second), accumulated with atomics
    int p_theta;
    double* aest;              // optional [n, M::SD]: REPORT(aest_all), nllk_ctcrw.hpp:246-249
    ScanDesc fdesc, bdesc;
    int ntiles;
    int summary;               // 1: stop after the tile prefixes are published (time-sharded runs only
                               // need the composite element of the whole shard = last inclusive prefix)
    int tile_lo;               // first ticket of this launch (> 0: summary over the tail of the shard only)
    int rerun;                 // forward kernel: 1 = phase (4), the filter re-run that yields the likelihood terms
                               // (and aest); 0 = stop at the checkpoints -- the adjoint kernel, which recomputes
                               // the forward states of every row anyway, sums the likelihood terms (llk_bwd)
    int llk_bwd;               // adjoint kernel: 1 = write tile_llk from its forward recomputation
};

// Start state of the track whose first row carries track index `idx` (stored, as a double, in
// the otherwise unused dt slot of track-start rows).
template <class M>
__device__ __forceinline__ typename M::State track_start_state(const KalmanArgs<typename M::R>& a, double idx) {
    return M::start_state(a.a0 + (size_t)idx * M::SD, a.P0);
}

// h = sigma_obs^2 = exp(2 log_sigma_obs), nllk_ctcrw.hpp:136,167
template <class R>
__device__ __forceinline__ R obs_variance(const KalmanArgs<R>& a) {
    return exp(2.0 * ScalarOf<R>::make(a.par[0], a.par_dot ? a.par_dot[0] : 0.0));
}

// the 8 row flags of this lane's chunk, packed (0xff = row beyond the end)
__device__ __forceinline__ unsigned long long load_flags8(const uint8_t* __restrict__ flags, int64_t base) {
    unsigned long long fl = 0;
#pragma unroll
    for (int k = 0; k < LC; ++k) fl |= (unsigned long long)flags[base + k * 32] << (8 * k);
    return fl;
}

// Per-plane pointers to this lane's first row of a warp-tile: row k of the lane is element k * 32
// of every plane, so a row costs one offset and one load per plane (no 64-bit index arithmetic
// per load).  `prefetch`: lanes 0 .. planes-1 ask for the warp-tile's 2 KB of their plane in L2, so
// that only the first of the LC row fetches pays the HBM latency.
template <class M>
struct RowPlanes {
    const typename M::R* wg[M::NW];
    const double* dt;
    const double* obs[M::ND];
    // tau, e, s2 (forward -> adjoint) are stored row-step by row-step, [n_pad / 32][NW][32]: one pointer and
    // constant offsets per row
    static constexpr int WGS = 32 * M::NW;          // wg: elements between two rows of a lane
};
template <class M>
__device__ __forceinline__ RowPlanes<M> open_planes(const KalmanArgs<typename M::R>& a, int64_t base, bool with_wg, bool prefetch) {
    RowPlanes<M> p;
    const int64_t np = a.X.n_pad;
#pragma unroll
    for (int c = 0; c < M::NW; ++c) {
        p.wg[c] = a.wg + (size_t)(base - (threadIdx.x & 31)) * M::NW + c * 32 + (threadIdx.x & 31);
    }
    p.dt = a.dt + base;
#pragma unroll
    for (int d = 0; d < M::ND; ++d) p.obs[d] = a.obs + (size_t)d * np + base;
    if (prefetch) {
        const int lane = threadIdx.x & 31;
        const int64_t q0 = base - lane;                       // first element of the warp-tile in every plane
        if (lane == 0) prefetch_l2(a.dt + q0, WT * 8);
        else if (lane <= M::ND) prefetch_l2(a.obs + (size_t)(lane - 1) * np + q0, WT * 8);
        else if (with_wg && lane <= M::ND + M::NW) prefetch_l2(a.wg + (size_t)q0 * M::NW + (size_t)(lane - 1 - M::ND) * WT, WT * (unsigned)sizeof(typename M::R));
    }
    return p;
}

// ---------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------
template <class M, int NT>
struct FwdSmem {
    using R = typename M::R;
    static constexpr int ES = (M::FwdElem::NDBL > 24 * ScalarOf<R>::NDBL) ? M::FwdElem::NDBL : 24 * ScalarOf<R>::NDBL;
    double stage[NT / 32][STAGE_DBL];
    double wagg[2][NT / 32][ES];     // shared scratch is double-buffered by tile parity: the only
    double tagg[2][ES];              // barrier between two tiles is the one that hands out the ticket
    R misc[2][16];
    R th[NT / 32][TH_CACHE];
    double thm[NT / 32][STAGE_SLOTS * MAX_NP];   // dense coefficient table of the warp-tile (R = double)
    uint64_t bar[NT / 32];
    int ticket[2];
    int early[2];                    // the tile aggregate was published before the barrier
};

template <class M, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) ctcrw_fwd_kernel(KalmanArgs<typename M::R> a) {
    using R = typename M::R;
    using SM = FwdSmem<M, NT>;
    using Ops = FwdOps<M>;
    using Elem = typename M::FwdElem;
    using St = typename M::State;
    constexpr int ND = M::ND;
    constexpr int NWARP = NT / 32;
    constexpr int NP = M::NP;
    static_assert(Elem::NDBL <= SM::ES, "element too large for the shared staging area");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const R h = obs_variance(a);                        // H = sigma_obs^2 I, nllk_ctcrw.hpp:136,167
    const bool mu0 = *a.mu_zero != 0;
    WarpStage st;
    stage_init(st, sm.stage[warp], &sm.bar[warp]);
    for (int i = lane; i < STAGE_DBL; i += 32) sm.stage[warp][i] = 0.0;     // row_eta_dense reads up to 3 slots past a row-step
    __syncwarp();
    mbar_fence_init();
    constexpr bool DENSE_ETA = SSDE_ETA_DENSE && std::is_same<R, double>::value;
    ThetaKey tkey;

    // Tickets are taken just in time (a tile reserved ahead of time by a busy CTA would stall the
    // look-back of every later tile).
    for (int it = 0;; ++it) {
        const int par = it & 1;
        if (tid == 0) sm.ticket[par] = a.tile_lo + (int)atomicAdd(a.fdesc.ticket, 1u);
        __syncthreads();
        const int tile = sm.ticket[par];
        if (tile >= a.ntiles) break;
#ifdef SSDE_STATS
        long long tc0 = clock64(), tc1 = 0, tc2 = 0, tc3 = 0;
#endif
        const int64_t q = (int64_t)tile * NWARP + warp;
        const int64_t base = q * WT + lane;
        const int64_t row0 = q * WT + (int64_t)lane * LC;
        // flags, dt and observations depend on the ticket only: request them before the descriptor chain
        const unsigned long long fl = load_flags8(a.flags, base);
        const RowPlanes<M> pl = open_planes<M>(a, base, false, SSDE_PLANE_PREFETCH != 0);
        double dt_nx = ((uint8_t)fl != 0xff) ? pl.dt[0] : 1.0, y_nx[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) y_nx[d] = ((uint8_t)fl != 0xff) ? pl.obs[d][0] : 0.0;
        const WtViewT<R> w = open_warptile<R>(a.X, q, a.theta, sm.th[warp], true, SSDE_THETA_REUSE ? &tkey : nullptr);
        if constexpr (DENSE_ETA) {
            if (w.staged && (!SSDE_THETA_REUSE || tkey.fresh)) fill_theta_matrix(w, sm.thm[warp]);
        }
        if (w.staged && lane == 0) {
            stage_issue(w, st, 0);
#if SSDE_FWD_PREFETCH == 1
            prefetch_l2(w.blk + (size_t)w.SV * 32, (unsigned)((LC - 1) * w.SV * 32 * 8));
#elif SSDE_FWD_PREFETCH == 2
            prefetch_l2(w.blk + (size_t)w.SV * 32, (unsigned)(2 * w.SV * 32 * 8));
#endif
        }

        // (1) thread element over its LC rows
        Elem E = M::fwd_identity();
        // dt and the observations of a row are fetched one row ahead of their use
        const unsigned step_bytes = (unsigned)w.SV * 32u * 8u;      // one row-step of design values
        const size_t step_dbl = (size_t)w.SV * 32;
        const double* next_step = w.blk + step_dbl;                 // row-step k + 1
#pragma unroll 1
        for (int k = 0; k < LC; ++k) {
            const int64_t pos = base + k * 32;
            const uint8_t f = (uint8_t)(fl >> (8 * k));
            const bool live = f != 0xff;
            const bool step = live && !(f & ROW_START);
            const double dtv = dt_nx;
            double y[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) y[d] = y_nx[d];
            if (k + 1 < LC) {
                dt_nx = pl.dt[(k + 1) * 32];                 // n_pad entries exist; a padding row's values are never used
#pragma unroll
                for (int d = 0; d < ND; ++d) y_nx[d] = pl.obs[d][(k + 1) * 32];
            }
            R eta[NP];
            if (w.staged) {
                stage_wait(st);
                if constexpr (DENSE_ETA) {
                    if (mu0) row_eta_dense<NP, ND>(w, st.buf, sm.thm[warp], eta);
                    else row_eta_dense<NP, 0>(w, st.buf, sm.thm[warp], eta);
                } else {
                    row_eta_staged<NP>(w, st, eta);
                }
                __syncwarp();
                if (lane == 0 && k + 1 < LC) {
                    stage_issue_at(st, next_step, step_bytes);
#if SSDE_FWD_PREFETCH == 2
                    if (k + 3 < LC) prefetch_l2(next_step + 2 * step_dbl, step_bytes);
#endif
                }
                next_step += step_dbl;
            } else if (step) {
                row_eta<NP>(w, k, a.theta, eta);
            }
            if (step) {
                const typename M::RowPar rp = M::transform(eta, dtv);
                M::store_rowpar(rp, [&](int c) -> R& { return const_cast<R&>(pl.wg[c][k * RowPlanes<M>::WGS]); });
                const typename M::Step sp = M::make_step(rp, dtv);
                M::fwd_append(E, sp, y, eta, (f & ROW_OBS) != 0, M::row_h(h, a.Hrow, a.X.n_pad, pos));
            } else if (live) {
                M::fwd_append_start(E, track_start_state<M>(a, dtv));
            }
        }
        // (2) warp inclusive scan (lower lanes = earlier rows)
        Elem inc = E;
#pragma unroll 1
        for (int o = 1; o < 32; o <<= 1) {
            Elem f = shfl_up_elem(inc, o);
            if (lane >= o) inc = M::fwd_combine(f, inc);
        }
        if (lane == 31) {
            store_elem(sm.wagg[par][warp], inc);
            if (warp == NWARP - 1) {
                // the last warp's rows alone usually form a constant map already: then they ARE the
                // tile aggregate and later tiles need not wait for this CTA's barrier
                const bool c = Ops::is_const(inc);
                sm.early[par] = c ? 1 : 0;
                if (c) { publish_agg<Ops>(a.fdesc, tile, inc); store_elem(sm.tagg[par], inc); }
            }
        }
        // The lane's exclusive prefix = its left neighbour's inclusive one.  The staging buffer is idle until
        // the next tile: every lane parks `inc` there and reads lane - 1's column after the look-back -- no
        // shuffle, and nothing for the register allocator to spill across the CTA barriers (it spilled the
        // element to local memory: 36 STL + 18 LDL.64 per thread and tile, ~1.7 GB of DRAM writes per launch
        // at 1e8 rows).
        constexpr bool PARK = Elem::NDBL * 32 <= STAGE_DBL;       // the element fits the warp's staging buffer
        Elem exc;
        if (PARK) {
            __syncwarp();                                          // every lane is done with the last row-step
            const double* ex = reinterpret_cast<const double*>(&inc);
#pragma unroll
            for (int i = 0; i < Elem::NDBL; ++i) st.buf[i * 32 + lane] = ex[i];
        } else {
            exc = shfl_up_elem(inc, 1);
            if (lane == 0) exc = M::fwd_identity();
        }
#ifdef SSDE_STATS
        tc1 = clock64();
#endif
        __syncthreads();
#ifdef SSDE_STATS
        tc2 = clock64();
#endif
        // (3) the last warp composes and publishes the tile aggregate while warp 0 looks back over
        //     the earlier tiles; warp 0 then publishes the inclusive prefix and the tile start state
        if (warp == NWARP - 1 && !sm.early[par]) {
            Elem tagg = load_elem<Elem>(sm.wagg[par][NWARP - 1]);
#pragma unroll 1
            for (int ww = NWARP - 2; ww >= 0 && !Ops::is_const(tagg); --ww) tagg = M::fwd_combine(load_elem<Elem>(sm.wagg[par][ww]), tagg);
            if (lane == 0) { publish_agg<Ops>(a.fdesc, tile, tagg); store_elem(sm.tagg[par], tagg); }
        }
        Elem pre;
        if (warp == 0) {
            pre = lookback<Ops>(a.fdesc, tile, a.tile_lo);
            if (lane == 0) {
                // state at the first row of the tile
                const St s0 = a.s_in ? M::load_state([&](int i) { return a.s_in[i]; }) : M::zero_state(a.P0);
                const St st0 = M::fwd_apply(pre, s0);
                M::store_state(st0, [&](int i) -> R& { return sm.misc[par][i]; });
            }
        }
        __syncthreads();
        if (warp == 0 && lane == 0) {
            // a constant-map aggregate already serves as the inclusive prefix (status code 3); the
            // summary pass of a time shard reads the last tile's inclusive element from f_incl
            const Elem tagg = load_elem<Elem>(sm.tagg[par]);
            if (a.summary || !Ops::is_const(tagg)) publish_incl<Ops>(a.fdesc, tile, M::fwd_combine(pre, tagg));
        }
#ifdef SSDE_STATS
        tc3 = clock64();
#endif
        if (a.summary) continue;
        // (4) exact start state of this thread, checkpoint, plain filter re-run
        // warp aggregates in front of this warp, starting at the nearest constant map (warp-uniform)
        St s = M::load_state([&](int i) { return sm.misc[par][i]; });
        {
            int w0 = warp - 1;
            while (w0 > 0 && !Ops::is_const(load_elem<Elem>(sm.wagg[par][w0]))) --w0;
#pragma unroll 1
            for (int ww = max(w0, 0); ww < warp; ++ww) s = M::fwd_apply(load_elem<Elem>(sm.wagg[par][ww]), s);
        }
        if (PARK) {
            Elem ex2;
            double* ex = reinterpret_cast<double*>(&ex2);
            const int src = lane ? lane - 1 : 0;
#pragma unroll
            for (int i = 0; i < Elem::NDBL; ++i) ex[i] = st.buf[i * 32 + src];
            if (lane == 0) ex2 = M::fwd_identity();
            s = M::fwd_apply(ex2, s);
        } else {
            s = M::fwd_apply(exc, s);
        }
        const int64_t chunk = q * 32 + lane;
        M::store_state(s, [&](int i) -> R& { return a.ckpt[(size_t)i * a.nchunks + chunk]; });
        if (!a.rerun) continue;        // the adjoint kernel sums the likelihood terms (llk_bwd)
        // sum_i log F_i is taken as the log of a running product (one log per chunk instead of
        // one per row); the product is folded into `slog` whenever it leaves a safe range.
        R quad = 0.0, fprod = 1.0, slog = 0.0;
        bool bad_f = false;
        dt_nx = ((uint8_t)fl != 0xff) ? pl.dt[0] : 1.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) y_nx[d] = ((uint8_t)fl != 0xff) ? pl.obs[d][0] : 0.0;
#pragma unroll 1
        for (int k = 0; k < LC; ++k) {
            const uint8_t f = (uint8_t)(fl >> (8 * k));
            if (f == 0xff) break;
            const int64_t pos = base + k * 32;
            const double dtv = dt_nx;
            double y[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) y[d] = y_nx[d];
            if (k + 1 < LC && (uint8_t)(fl >> (8 * (k + 1))) != 0xff) {
                dt_nx = pl.dt[(k + 1) * 32];
#pragma unroll
                for (int d = 0; d < ND; ++d) y_nx[d] = pl.obs[d][(k + 1) * 32];
            }
            if (f & ROW_START) {
                s = track_start_state<M>(a, dtv);
            } else {
                R mu[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) mu[d] = 0.0;
                if (!mu0) row_eta_prefix<ND>(w, k, a.theta, mu);
                // the row's step, rebuilt from the tau, e, s2 the row loop has just written (an L1 / L2 hit): this
                // phase only runs when no adjoint pass follows, and a 40 KB step cache in shared memory would cost
                // the kernel its fourth CTA per SM
                const typename M::Step sp = M::make_step(M::load_rowpar([&](int c) { return pl.wg[c][k * RowPlanes<M>::WGS]; }), dtv);
                R F, qd;
                M::template fwd_step<false>(s, sp, y, mu, (f & ROW_OBS) != 0, M::row_h(h, a.Hrow, a.X.n_pad, pos), nullptr, F, qd);
                quad += qd;
                fprod *= F;
                bad_f |= value(F) <= 0.0;          // the reference's detF <= 0 branch (nllk_ctcrw.hpp:226-228) is not built: flag it
                if (!(value(fprod) > 1e-150 && value(fprod) < 1e150)) { slog += log(fprod); fprod = 1.0; }
            }
            if (a.aest) {
                M::store_mean(s, a.aest + (size_t)(row0 + k) * M::SD);
            }
        }
        const double llk = warp_sum(value(-0.5 * ((double)M::LOGF_MULT * (slog + log(fprod)) + quad)));
        if (lane == 0) a.tile_llk[q] = llk;              // one partial per warp-tile
        if (bad_f) atomicOr(a.fdesc.error, 2u);
#ifdef SSDE_STATS
        if (lane == 0) {                                 // per-warp phase cycles: 4 + 4*warp-class (warp 0 / others)
            unsigned long long* st_ = a.fdesc.stats + 4 + (warp == 0 ? 0 : 4);
            atomicAdd(st_ + 0, (unsigned long long)(tc1 - tc0));        // phase 1-2 (elements + warp scan)
            atomicAdd(st_ + 1, (unsigned long long)(tc2 - tc1));        // first barrier
            atomicAdd(st_ + 2, (unsigned long long)(tc3 - tc2));        // aggregate / look-back / second barrier
            atomicAdd(st_ + 3, (unsigned long long)(clock64() - tc3));  // phase 4 (re-run)
        }
#endif
    }
}

// ---------------------------------------------------------------------------------------------
// adjoint kernel
// ---------------------------------------------------------------------------------------------
constexpr int SGRAD = 256;         // per-CTA gradient accumulators (doubles) when p_theta fits

template <class M, int NT>
struct BwdSmem {
    using R = typename M::R;
    // per row: the forward state before the row (M::FS scalars), later overwritten by eta_bar
    // (slots 0..NP-1) and by the transposed-product scratch (slots NP..RS-1)
    static constexpr int RS = (M::FS > M::NP + 3) ? M::FS : M::NP + 3;
    static constexpr int ES = (M::BwdElem::NDBL > 16 * ScalarOf<R>::NDBL) ? M::BwdElem::NDBL : 16 * ScalarOf<R>::NDBL;
    R Rs[LC][RS][NT];
    double wagg[2][NT / 32][ES];
    double tagg[2][ES];
    R misc[2][16];
    R th[NT / 32][TH_CACHE];
    R sgrad[SGRAD];
    int ticket[2];
    int early[2];
};

// per-row inputs of the adjoint sweep, fetched one row ahead
template <class M>
struct RowIn {
    typename M::RowPar rp;
    double dt, y[M::ND];
};
template <class M>
__device__ __forceinline__ RowIn<M> load_row(const RowPlanes<M>& p, int k, bool live) {
    RowIn<M> r;
    const int o = k * 32;
    // every per-row array is n_pad long and nothing loaded for a row that is not a filter step is ever
    // used (the callers branch on the row's flags), so the loads need no predicate
    (void)live;
    r.dt = p.dt[o];
    r.rp = M::load_rowpar([&](int c) { return p.wg[c][k * RowPlanes<M>::WGS]; });
#pragma unroll
    for (int d = 0; d < M::ND; ++d) r.y[d] = p.obs[d][o];
    return r;
}

template <class M, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) ctcrw_bwd_kernel(KalmanArgs<typename M::R> a) {
    using R = typename M::R;
    using SM = BwdSmem<M, NT>;
    using Ops = BwdOps<M>;
    using Elem = typename M::BwdElem;
    using St = typename M::State;
    using Ad = typename M::Adj;
    constexpr int ND = M::ND;
    constexpr int NWARP = NT / 32;
    constexpr int NP = M::NP;
    constexpr int FS = SM::RS;
    constexpr int TCAP = (FS - NP) * LC;          // slots whose scratch fits in the spare row slots
    static_assert(Elem::NDBL <= SM::ES, "element too large for the shared staging area");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const R h = obs_variance(a);
    const bool mu0 = *a.mu_zero != 0;
    GradAccT<R> gacc{(a.p_theta <= SGRAD) ? sm.sgrad : nullptr, a.grad_theta, a.p_theta};
    if (gacc.sgrad) for (int i = tid; i < SGRAD; i += NT) sm.sgrad[i] = 0.0;

    for (int it = 0;; ++it) {
        const int par = it & 1;
        if (tid == 0) sm.ticket[par] = a.tile_lo + (int)atomicAdd(a.bdesc.ticket, 1u);
        __syncthreads();
        const int ticket = sm.ticket[par];
        if (ticket >= a.ntiles) break;
        const int tile = a.ntiles - 1 - ticket;           // reverse time order
#ifdef SSDE_STATS
        long long tc0 = clock64(), tc1 = 0, tc2 = 0, tc3 = 0;
#endif
        const int64_t q = (int64_t)tile * NWARP + warp;
        const int64_t base = q * WT + lane;
        const int64_t chunk = q * 32 + lane;
        const WtViewT<R> w = open_warptile<R>(a.X, q, a.theta, sm.th[warp], !mu0);
        const unsigned long long fl = load_flags8(a.flags, base);
#if SSDE_BWD_PREFETCH == 2
        if (lane == 0 && !a.summary) prefetch_l2(w.blk, (unsigned)(LC * w.SV * 32 * 8));
#endif

        // (1) recompute the forward states of this thread's rows from its checkpoint and compose
        //     the rows' adjoint elements (in time order)
        St s = M::load_state([&](int i) { return a.ckpt[(size_t)i * a.nchunks + chunk]; });
        Elem E = M::bwd_identity();
        const RowPlanes<M> pl = open_planes<M>(a, base, true, SSDE_PLANE_PREFETCH != 0);
        RowIn<M> nx = load_row<M>(pl, 0, (uint8_t)fl != 0xff);
        // llk_bwd: this recomputation IS the filter over the rows -- F and the quadratic form of every
        // row come out of it, so the likelihood terms are summed here (same running-product form as the
        // forward kernel's re-run, which is then skipped)
        const bool want_llk = a.llk_bwd && !a.summary;
        R quad = 0.0, fprod = 1.0, slog = 0.0;
        bool bad_f = false;
#pragma unroll 1
        for (int k = 0; k < LC; ++k) {
            const uint8_t f = (uint8_t)(fl >> (8 * k));
            const bool live = f != 0xff;
            const bool step = live && !(f & ROW_START);
            const RowIn<M> r = nx;
            if (k + 1 < LC) nx = load_row<M>(pl, k + 1, (uint8_t)(fl >> (8 * (k + 1))) != 0xff);
            // state BEFORE row k
            M::store_state(s, [&](int i) -> R& { return sm.Rs[k][i][tid]; });
            if (step) {
                R mu[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) mu[d] = 0.0;
                if (!mu0) row_eta_prefix<ND>(w, k, a.theta, mu);
                const typename M::Step sp = M::make_step(r.rp, r.dt);
                typename M::Aux ax;
                R F, qd;
                M::template fwd_step<true>(s, sp, r.y, mu, (f & ROW_OBS) != 0, M::row_h(h, a.Hrow, a.X.n_pad, base + k * 32), &ax, F, qd);
                E = M::bwd_append_row(E, sp, ax, (f & ROW_OBS) != 0, (f & ROW_LAST) != 0);
                if (want_llk) {
                    quad += qd;
                    fprod *= F;
                    bad_f |= value(F) <= 0.0;          // nllk_ctcrw.hpp:226-228 is not built: flag it
                    if (!(value(fprod) > 1e-150 && value(fprod) < 1e150)) { slog += log(fprod); fprod = 1.0; }
                }
            } else if (live) {
                s = track_start_state<M>(a, r.dt);
                E = M::bwd_combine(E, M::bwd_const(M::adj_zero()));
            }
        }
        if (want_llk) {
            const double llk = warp_sum(value(-0.5 * ((double)M::LOGF_MULT * (slog + log(fprod)) + quad)));
            if (lane == 0) a.tile_llk[q] = llk;              // one partial per warp-tile
            if (bad_f) atomicOr(a.bdesc.error, 2u);
        }
        // (2) warp inclusive SUFFIX scan (higher lanes = later rows)
        Elem inc = E;
#pragma unroll 1
        for (int o = 1; o < 32; o <<= 1) {
            Elem f = shfl_down_elem(inc, o);
            if (lane + o < 32) inc = M::bwd_combine(inc, f);
        }
        if (lane == 0) {
            store_elem(sm.wagg[par][warp], inc);
            if (warp == 0) {                     // the earliest rows decide (adjoint flows backwards in time)
                const bool c = Ops::is_const(inc);
                sm.early[par] = c ? 1 : 0;
                if (c) { publish_agg<Ops>(a.bdesc, ticket, inc); store_elem(sm.tagg[par], inc); }
            }
        }
        Elem exc = shfl_down_elem(inc, 1);
        if (lane == 31) exc = M::bwd_identity();
#ifdef SSDE_STATS
        tc1 = clock64();
#endif
        __syncthreads();
#ifdef SSDE_STATS
        tc2 = clock64();
#endif
        // (3) the last warp composes and publishes the tile aggregate while warp 0 looks back over
        //     the LATER tiles; warp 0 then publishes the inclusive suffix and the adjoint entering
        //     the tile end
        if (warp == NWARP - 1 && !sm.early[par]) {
            // the adjoint flows from later rows to earlier ones: start with the EARLIEST warp
            Elem tagg = load_elem<Elem>(sm.wagg[par][0]);
#pragma unroll 1
            for (int ww = 1; ww < NWARP && !Ops::is_const(tagg); ++ww) tagg = M::bwd_combine(tagg, load_elem<Elem>(sm.wagg[par][ww]));
            if (lane == 0) { publish_agg<Ops>(a.bdesc, ticket, tagg); store_elem(sm.tagg[par], tagg); }
        }
        Elem suf;
        if (warp == 0) {
            suf = lookback<Ops>(a.bdesc, ticket, a.tile_lo);
            if (lane == 0) {
                const Ad g0 = a.g_in ? M::load_adj([&](int i) { return a.g_in[i]; }) : M::adj_zero();
                const Ad gt = M::bwd_apply(suf, g0);
                M::store_adj(gt, [&](int i) -> R& { return sm.misc[par][i]; });
            }
        }
        __syncthreads();
        if (warp == 0 && lane == 0) {
            const Elem tagg = load_elem<Elem>(sm.tagg[par]);
            if (a.summary || !Ops::is_const(tagg)) publish_incl<Ops>(a.bdesc, ticket, M::bwd_combine(tagg, suf));
        }
#ifdef SSDE_STATS
        tc3 = clock64();
#endif
        if (a.summary) continue;
#if SSDE_BWD_PREFETCH == 1
        // the transposed design product at the end of the tile streams LC x S x 32 values of this
        // warp-tile from HBM: ask for them now so that they wait in L2 when the sweep is done
        if (lane == 0) prefetch_l2(w.blk, (unsigned)(LC * w.SV * 32 * 8));
#endif
        // (4) adjoint entering this thread's last row, then the reverse sweep over its rows
        Ad g = M::load_adj([&](int i) { return sm.misc[par][i]; });
        {
            int w1 = warp + 1;
            while (w1 < NWARP - 1 && !Ops::is_const(load_elem<Elem>(sm.wagg[par][w1]))) ++w1;
#pragma unroll 1
            for (int ww = min(w1, NWARP - 1); ww > warp; --ww) g = M::bwd_apply(load_elem<Elem>(sm.wagg[par][ww]), g);
        }
        g = M::bwd_apply(exc, g);
        R gh = 0.0;
        nx = load_row<M>(pl, LC - 1, (uint8_t)(fl >> (8 * (LC - 1))) != 0xff);
#pragma unroll 1
        for (int k = LC - 1; k >= 0; --k) {
            const uint8_t f = (uint8_t)(fl >> (8 * k));
            const RowIn<M> r = nx;
            if (k > 0) nx = load_row<M>(pl, k - 1, (uint8_t)(fl >> (8 * (k - 1))) != 0xff);
            R gp[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) gp[j] = 0.0;
            if (f != 0xff) {
                if (f & ROW_START) {
                    g = M::adj_zero();
                } else {
                    R mu[ND];
#pragma unroll
                    for (int d = 0; d < ND; ++d) mu[d] = 0.0;
                    if (!mu0) row_eta_prefix<ND>(w, k, a.theta, mu);
                    St sk = M::load_state([&](int i) { return sm.Rs[k][i][tid]; });
                    const typename M::Step sp = M::make_step(r.rp, r.dt);
                    const bool has = (f & ROW_OBS) != 0, cut = (f & ROW_LAST) != 0;
                    typename M::Aux ax;
                    R F, qd;
                    const typename M::Hc hr = M::row_h(h, a.Hrow, a.X.n_pad, base + k * 32);
                    M::template fwd_step<true>(sk, sp, r.y, mu, has, hr, &ax, F, qd);
                    const Ad gin = cut ? M::adj_zero() : g;
                    R g_h;
                    M::row_param_grad(gin, sp, ax, mu, r.rp, r.dt, has, hr, gp, g_h);
                    gh += g_h;
                    g = M::bwd_apply_row(sp, ax, has, cut, g);
                }
            }
            // eta_bar of this row replaces its (consumed) forward state
#pragma unroll
            for (int j = 0; j < NP; ++j) sm.Rs[k][j][tid] = gp[j];
        }
#ifdef SSDE_STATS
        const long long tc35 = clock64();
#endif
        // (5) grad_theta += X' eta_bar for this warp-tile
        if (w.uniform && w.S <= TCAP) {
            scatter_warptile_transposed<NP, R>(w, gacc, [&](int k, int p) { return sm.Rs[k][p][tid]; },
                                               [&](int j, int l) -> R& { return sm.Rs[j / (FS - NP)][NP + j % (FS - NP)][(tid & ~31) + l]; });
        } else {
            scatter_warptile<NP, R>(w, gacc, [&](int k, int p) { return sm.Rs[k][p][tid]; });
        }
#ifdef SSDE_STATS
        const long long tc4 = clock64();
#endif
        gh = warp_sum(gh);
#ifdef SSDE_STATS
        if (lane == 0) {
            unsigned long long* st_ = a.bdesc.stats + 4 + (warp == 0 ? 0 : 4);
            atomicAdd(st_ + 0, (unsigned long long)(tc1 - tc0));
            atomicAdd(st_ + 1, (unsigned long long)(tc2 - tc1));
            atomicAdd(st_ + 2, (unsigned long long)(tc3 - tc2));
            atomicAdd(st_ + 3, (unsigned long long)(tc35 - tc3));
            atomicAdd(a.bdesc.stats + 12 + (warp == 0 ? 0 : 1), (unsigned long long)(tc4 - tc35));
        }
#endif
        if (lane == 0) {                                 // one partial per warp-tile
            a.tile_gh[q] = value(gh);
            if constexpr (!std::is_same<R, double>::value) a.tile_gh[a.X.n_pad / WT + q] = gh.d;
        }
    }
    if (gacc.sgrad) {
        __syncthreads();
        grad_flush(gacc, NT);
    }
}

// ---------------------------------------------------------------------------------------------
// time-sharded runs: composite elements of the shards (gathered from all ranks) -> incoming
// state / adjoint of shard `me`.  One thread.
// ---------------------------------------------------------------------------------------------
// Shard element handed to the other ranks: the inclusive prefix of the shard's last ticket plus
// one double that says whether it is a constant map.  A summary pass over the TAIL of a shard is
// the shard's exact composite element iff that flag is set (nothing before the tail matters);
// otherwise the host repeats the pass over the whole shard.
template <class Elem, class Ops>
__global__ void shard_elem_kernel(const double* __restrict__ incl_last, double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const Elem e = load_elem<Elem>(incl_last);
    store_elem(out, e);
    out[Elem::NDBL] = Ops::is_const(e) ? 1.0 : 0.0;
}

template <class M>
__global__ void shard_state_kernel(const double* __restrict__ elems, int n_shards, int me, PriorCov P0,
                                   typename M::R* __restrict__ s_out) {
    using R = typename M::R;
    using Elem = typename M::FwdElem;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Elem acc = M::fwd_identity();
    for (int i = 0; i < me && i < n_shards; ++i) acc = M::fwd_combine(acc, load_elem<Elem>(elems + (size_t)i * (Elem::NDBL + 1)));
    const typename M::State s = M::fwd_apply(acc, M::zero_state(P0));   // shard 0 begins with a track start: the input is irrelevant
    M::store_state(s, [&](int i) -> R& { return s_out[i]; });
}

template <class M>
__global__ void shard_adjoint_kernel(const double* __restrict__ elems, int n_shards, int me,
                                     typename M::R* __restrict__ g_out) {
    using R = typename M::R;
    using Elem = typename M::BwdElem;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Elem acc = M::bwd_identity();
    for (int i = n_shards - 1; i > me; --i) acc = M::bwd_combine(load_elem<Elem>(elems + (size_t)i * (Elem::NDBL + 1)), acc);
    const typename M::Adj g = M::bwd_apply(acc, M::adj_zero());
    M::store_adj(g, [&](int i) -> R& { return g_out[i]; });
}

}  // namespace ssde
