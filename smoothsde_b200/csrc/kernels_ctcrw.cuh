// CTCRW Kalman filter as a time-parallel scan: forward (likelihood) and adjoint (gradient)
// kernels.  Replaces the sequential loop nllk_ctcrw.hpp:195-247 and TMB's reverse sweep of it.
//
// Rows of all tracks are stacked exactly as in the reference's data list; a track start is a
// "constant map" element, so the whole stack is ONE segmented scan and the same kernels serve
// many short tracks, few long tracks and a single 1e8-row track.  Work decomposition:
//   tile = NT threads x LC consecutive rows per thread, tiles taken from a dynamic ticket,
//   (1) each thread composes its LC rows into one element (fwd_append),
//   (2) warp shuffle scan + cross-warp scan of the thread elements,
//   (3) chained look-back across tiles gives the state at the tile start,
//   (4) each thread re-runs the plain filter over its rows from its exact start state.
// The adjoint kernel walks the tiles in reverse with elements (L, z, D).
#pragma once

#include "common.cuh"

namespace ssde {

template <int ND>
struct CtcrwArgs {
    int64_t n;                 // rows in this shard
    int ntiles;
    const double* W;           // [n, ND+3]: mu_1..mu_ND, tau, e, s2   (written by linpred kernel)
    const double* obs;         // [n, ND] row-major, NA replaced by 0
    const double* dt;          // [n]
    const uint8_t* flags;      // [n]
    const int64_t* track_starts;   // [n_tracks] sorted rows flagged ROW_START
    const double* a0;          // [n_tracks, 2*ND]
    int n_tracks;
    Sym2 P0;
    const double* par;         // device parameter vector; par[0] = log_sigma_obs
    const double* s_in;        // optional incoming state (2*ND + 3 doubles) for a continued shard
    const double* g_in;        // optional incoming adjoint (2*ND + 3 doubles)
    double* ckpt;              // [(2*ND+3), nchunks] start state of every thread chunk
    int64_t nchunks;
    double* tile_llk;          // [ntiles]
    double* tile_gh;           // [ntiles]
    double* eta_bar;           // [n, ND+2] adjoint of the linear predictors
    double* aest;              // optional [n, 2*ND]: REPORT(aest_all), nllk_ctcrw.hpp:246-249
    ScanDesc fdesc, bdesc;
};

template <int ND>
__device__ __forceinline__ State<ND> track_start_state(const CtcrwArgs<ND>& a, int64_t row) {
    // binary search: index of `row` in track_starts
    int lo = 0, hi = a.n_tracks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.track_starts[mid] <= row) lo = mid; else hi = mid - 1;
    }
    State<ND> s;
    const double* p = a.a0 + (size_t)lo * 2 * ND;
#pragma unroll
    for (int d = 0; d < ND; ++d) s.a[d] = {p[2 * d], p[2 * d + 1]};
    s.P = a.P0;
    return s;
}

template <int ND>
__device__ __forceinline__ State<ND> load_state(const double* p) {
    State<ND> s;
#pragma unroll
    for (int d = 0; d < ND; ++d) s.a[d] = {p[2 * d], p[2 * d + 1]};
    s.P = {p[2 * ND], p[2 * ND + 1], p[2 * ND + 2]};
    return s;
}

template <int ND, int NT, int LC>
struct CtcrwSmem {
    static constexpr int NW = ND + 3;
    using SW = Staged<NW, NT, LC>;
    using SY = Staged<ND, NT, LC>;
    using SD = Staged<1, NT, LC>;
    static constexpr int FS = 2 * ND + 3;               // doubles of a forward state
    using SF = Staged<FS, NT, LC>;
    static constexpr int OFF_W = 0;
    static constexpr int OFF_Y = OFF_W + SW::SIZE;
    static constexpr int OFF_DT = OFF_Y + SY::SIZE;
    static constexpr int OFF_WAGG = OFF_DT + SD::SIZE;   // NT/32 elements (<= 24 doubles each)
    static constexpr int OFF_MISC = OFF_WAGG + (NT / 32) * 24;
    static constexpr int OFF_FS = OFF_MISC + 32;          // backward only
    static constexpr int DBL_FWD = OFF_FS;
    static constexpr int DBL_BWD = OFF_FS + SF::SIZE;
    static constexpr size_t BYTES_FWD = (size_t)DBL_FWD * 8 + (size_t)LC * (NT + 4);
    static constexpr size_t BYTES_BWD = (size_t)DBL_BWD * 8 + (size_t)LC * (NT + 4);
};

// ---------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------
template <int ND, int NT, int LC>
__global__ void __launch_bounds__(NT) ctcrw_fwd_kernel(CtcrwArgs<ND> a) {
    using SM = CtcrwSmem<ND, NT, LC>;
    using Ops = FwdOps<ND>;
    using Elem = FwdElem<ND>;
    constexpr int NWARP = NT / 32;
    static_assert(Elem::NDBL <= 24, "element too large for the shared staging area");
    extern __shared__ __align__(16) double smem[];
    double* sW = smem + SM::OFF_W;
    double* sY = smem + SM::OFF_Y;
    double* sDt = smem + SM::OFF_DT;
    double* sWagg = smem + SM::OFF_WAGG;
    double* sMisc = smem + SM::OFF_MISC;
    uint8_t* sFl = reinterpret_cast<uint8_t*>(smem + SM::DBL_FWD);
    __shared__ int s_ticket;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double h = exp(2.0 * a.par[0]);               // H = sigma_obs^2 I, nllk_ctcrw.hpp:136,167

    while (true) {
        __syncthreads();
        if (tid == 0) s_ticket = (int)atomicAdd(a.fdesc.ticket, 1u);
        __syncthreads();
        const int tile = s_ticket;
        if (tile >= a.ntiles) break;
        const int64_t r0 = (int64_t)tile * (NT * LC);
        stage_rows<SM::NW, NT, LC>(sW, a.W, r0, a.n);
        stage_rows<ND, NT, LC>(sY, a.obs, r0, a.n);
        stage_rows<1, NT, LC>(sDt, a.dt, r0, a.n);
        stage_flags<NT, LC>(sFl, a.flags, r0, a.n);
        __syncthreads();

        // (1) thread element over its LC rows
        const int64_t row0 = r0 + (int64_t)tid * LC;
        Elem E = fwd_identity<ND>();
#pragma unroll 1
        for (int k = 0; k < LC; ++k) {
            const uint8_t f = sFl[k * (NT + 4) + tid];
            if (f == 0xff) break;
            if (f & ROW_START) {
                fwd_append_start<ND>(E, track_start_state<ND>(a, row0 + k));
            } else {
                double mu[ND], y[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    mu[d] = sW[SM::SW::at(k, d, tid)];
                    y[d] = sY[SM::SY::at(k, d, tid)];
                }
                const StepPar sp = make_step(sW[SM::SW::at(k, ND, tid)], sW[SM::SW::at(k, ND + 1, tid)],
                                             sW[SM::SW::at(k, ND + 2, tid)], sDt[SM::SD::at(k, 0, tid)]);
                fwd_append<ND>(E, sp, y, mu, (f & ROW_OBS) != 0, h);
            }
        }
        // (2) warp inclusive scan (lower lanes = earlier rows)
        Elem inc = E;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            Elem f = shfl_up_elem(inc, o);
            if (lane >= o) inc = fwd_combine<ND>(f, inc);
        }
        if (lane == 31) store_elem(sWagg + warp * 24, inc);
        Elem exc = shfl_up_elem(inc, 1);
        if (lane == 0) exc = fwd_identity<ND>();
        __syncthreads();
        // (3) tile aggregate, chained look-back (warp 0), tile start state
        if (warp == 0) {
            Elem tagg = load_elem<Elem>(sWagg);
#pragma unroll
            for (int w = 1; w < NWARP; ++w) tagg = fwd_combine<ND>(tagg, load_elem<Elem>(sWagg + w * 24));
            if (lane == 0) publish_agg<Ops>(a.fdesc, tile, tagg);
            const Elem pre = lookback<Ops>(a.fdesc, tile);
            if (lane == 0) {
                publish_incl<Ops>(a.fdesc, tile, fwd_combine<ND>(pre, tagg));
                // state at the first row of the tile
                State<ND> s0;
                if (a.s_in) s0 = load_state<ND>(a.s_in);
                else { s0.P = a.P0;
#pragma unroll
                    for (int d = 0; d < ND; ++d) s0.a[d] = {0.0, 0.0}; }
                const State<ND> st = fwd_apply<ND>(pre, s0);
#pragma unroll
                for (int d = 0; d < ND; ++d) { sMisc[2 * d] = st.a[d].x; sMisc[2 * d + 1] = st.a[d].y; }
                sMisc[2 * ND] = st.P.a; sMisc[2 * ND + 1] = st.P.b; sMisc[2 * ND + 2] = st.P.c;
            }
        }
        __syncthreads();
        // (4) exact start state of this thread, checkpoint, plain filter re-run
        State<ND> s = load_state<ND>(sMisc);
#pragma unroll 1
        for (int w = 0; w < warp; ++w) s = fwd_apply<ND>(load_elem<Elem>(sWagg + w * 24), s);
        s = fwd_apply<ND>(exc, s);
        const int64_t chunk = (int64_t)tile * NT + tid;
        if (row0 < a.n) {
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                a.ckpt[(size_t)(2 * d) * a.nchunks + chunk] = s.a[d].x;
                a.ckpt[(size_t)(2 * d + 1) * a.nchunks + chunk] = s.a[d].y;
            }
            a.ckpt[(size_t)(2 * ND) * a.nchunks + chunk] = s.P.a;
            a.ckpt[(size_t)(2 * ND + 1) * a.nchunks + chunk] = s.P.b;
            a.ckpt[(size_t)(2 * ND + 2) * a.nchunks + chunk] = s.P.c;
        }
        double llk = 0.0;
#pragma unroll 1
        for (int k = 0; k < LC; ++k) {
            const uint8_t f = sFl[k * (NT + 4) + tid];
            if (f == 0xff) break;
            if (f & ROW_START) {
                s = track_start_state<ND>(a, row0 + k);
            } else {
                double mu[ND], y[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    mu[d] = sW[SM::SW::at(k, d, tid)];
                    y[d] = sY[SM::SY::at(k, d, tid)];
                }
                const StepPar sp = make_step(sW[SM::SW::at(k, ND, tid)], sW[SM::SW::at(k, ND + 1, tid)],
                                             sW[SM::SW::at(k, ND + 2, tid)], sDt[SM::SD::at(k, 0, tid)]);
                llk += fwd_step<ND, false>(s, sp, y, mu, (f & ROW_OBS) != 0, h, nullptr);
            }
            if (a.aest) {
                double* o = a.aest + (size_t)(row0 + k) * (2 * ND);
#pragma unroll
                for (int d = 0; d < ND; ++d) { o[2 * d] = s.a[d].x; o[2 * d + 1] = s.a[d].y; }
            }
        }
        const double tl = block_sum<NT>(llk, sMisc + 16);
        if (tid == 0) a.tile_llk[tile] = tl;
    }
}

// ---------------------------------------------------------------------------------------------
// adjoint kernel
// ---------------------------------------------------------------------------------------------
template <int ND>
__device__ __forceinline__ Adj<ND> load_adj(const double* p) {
    Adj<ND> g;
#pragma unroll
    for (int d = 0; d < ND; ++d) g.a[d] = {p[2 * d], p[2 * d + 1]};
    g.P = {p[2 * ND], p[2 * ND + 1], p[2 * ND + 2]};
    return g;
}

template <int ND, int NT, int LC>
__global__ void __launch_bounds__(NT) ctcrw_bwd_kernel(CtcrwArgs<ND> a) {
    using SM = CtcrwSmem<ND, NT, LC>;
    using Ops = BwdOps<ND>;
    using Elem = BwdElem<ND>;
    constexpr int NWARP = NT / 32;
    constexpr int NP = ND + 2;
    extern __shared__ __align__(16) double smem[];
    double* sW = smem + SM::OFF_W;
    double* sY = smem + SM::OFF_Y;
    double* sDt = smem + SM::OFF_DT;
    double* sWagg = smem + SM::OFF_WAGG;
    double* sMisc = smem + SM::OFF_MISC;
    double* sFs = smem + SM::OFF_FS;
    uint8_t* sFl = reinterpret_cast<uint8_t*>(smem + SM::DBL_BWD);
    __shared__ int s_ticket;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double h = exp(2.0 * a.par[0]);

    while (true) {
        __syncthreads();
        if (tid == 0) s_ticket = (int)atomicAdd(a.bdesc.ticket, 1u);
        __syncthreads();
        const int ticket = s_ticket;
        if (ticket >= a.ntiles) break;
        const int tile = a.ntiles - 1 - ticket;           // reverse time order
        const int64_t r0 = (int64_t)tile * (NT * LC);
        stage_rows<SM::NW, NT, LC>(sW, a.W, r0, a.n);
        stage_rows<ND, NT, LC>(sY, a.obs, r0, a.n);
        stage_rows<1, NT, LC>(sDt, a.dt, r0, a.n);
        stage_flags<NT, LC>(sFl, a.flags, r0, a.n);
        __syncthreads();

        // (1) recompute the forward states of this thread's rows from its checkpoint and compose
        //     the rows' adjoint elements (in time order)
        const int64_t row0 = r0 + (int64_t)tid * LC;
        const int64_t chunk = (int64_t)tile * NT + tid;
        State<ND> s;
        if (row0 < a.n) {
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                s.a[d].x = a.ckpt[(size_t)(2 * d) * a.nchunks + chunk];
                s.a[d].y = a.ckpt[(size_t)(2 * d + 1) * a.nchunks + chunk];
            }
            s.P.a = a.ckpt[(size_t)(2 * ND) * a.nchunks + chunk];
            s.P.b = a.ckpt[(size_t)(2 * ND + 1) * a.nchunks + chunk];
            s.P.c = a.ckpt[(size_t)(2 * ND + 2) * a.nchunks + chunk];
        } else {
            s.P = a.P0;
#pragma unroll
            for (int d = 0; d < ND; ++d) s.a[d] = {0.0, 0.0};
        }
        Elem E = bwd_identity<ND>();
#pragma unroll 1
        for (int k = 0; k < LC; ++k) {
            const uint8_t f = sFl[k * (NT + 4) + tid];
            if (f == 0xff) break;
            // state BEFORE row k
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                sFs[SM::SF::at(k, 2 * d, tid)] = s.a[d].x;
                sFs[SM::SF::at(k, 2 * d + 1, tid)] = s.a[d].y;
            }
            sFs[SM::SF::at(k, 2 * ND, tid)] = s.P.a;
            sFs[SM::SF::at(k, 2 * ND + 1, tid)] = s.P.b;
            sFs[SM::SF::at(k, 2 * ND + 2, tid)] = s.P.c;
            if (f & ROW_START) {
                s = track_start_state<ND>(a, row0 + k);
                E = bwd_combine<ND>(E, bwd_const<ND>(adj_zero<ND>()));
            } else {
                double mu[ND], y[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    mu[d] = sW[SM::SW::at(k, d, tid)];
                    y[d] = sY[SM::SY::at(k, d, tid)];
                }
                const StepPar sp = make_step(sW[SM::SW::at(k, ND, tid)], sW[SM::SW::at(k, ND + 1, tid)],
                                             sW[SM::SW::at(k, ND + 2, tid)], sDt[SM::SD::at(k, 0, tid)]);
                StepAux<ND> ax;
                fwd_step<ND, true>(s, sp, y, mu, (f & ROW_OBS) != 0, h, &ax);
                E = bwd_combine<ND>(E, bwd_row_elem<ND>(sp, ax, (f & ROW_OBS) != 0, (f & ROW_LAST) != 0));
            }
        }
        // (2) warp inclusive SUFFIX scan (higher lanes = later rows)
        Elem inc = E;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            Elem f = shfl_down_elem(inc, o);
            if (lane + o < 32) inc = bwd_combine<ND>(inc, f);
        }
        if (lane == 0) store_elem(sWagg + warp * 24, inc);
        Elem exc = shfl_down_elem(inc, 1);
        if (lane == 31) exc = bwd_identity<ND>();
        __syncthreads();
        // (3) tile aggregate, chained look-back over LATER tiles, adjoint entering the tile end
        if (warp == 0) {
            Elem tagg = load_elem<Elem>(sWagg + (NWARP - 1) * 24);
#pragma unroll
            for (int w = NWARP - 2; w >= 0; --w) tagg = bwd_combine<ND>(load_elem<Elem>(sWagg + w * 24), tagg);
            if (lane == 0) publish_agg<Ops>(a.bdesc, ticket, tagg);
            const Elem suf = lookback<Ops>(a.bdesc, ticket);
            if (lane == 0) {
                publish_incl<Ops>(a.bdesc, ticket, bwd_combine<ND>(tagg, suf));
                Adj<ND> g0 = a.g_in ? load_adj<ND>(a.g_in) : adj_zero<ND>();
                const Adj<ND> gt = bwd_apply<ND>(suf, g0);
#pragma unroll
                for (int d = 0; d < ND; ++d) { sMisc[2 * d] = gt.a[d].x; sMisc[2 * d + 1] = gt.a[d].y; }
                sMisc[2 * ND] = gt.P.a; sMisc[2 * ND + 1] = gt.P.b; sMisc[2 * ND + 2] = gt.P.c;
            }
        }
        __syncthreads();
        // (4) adjoint entering this thread's last row, then the reverse sweep over its rows
        Adj<ND> g = load_adj<ND>(sMisc);
#pragma unroll 1
        for (int w = NWARP - 1; w > warp; --w) g = bwd_apply<ND>(load_elem<Elem>(sWagg + w * 24), g);
        g = bwd_apply<ND>(exc, g);
        double gh = 0.0;
#pragma unroll 1
        for (int k = LC - 1; k >= 0; --k) {
            const uint8_t f = sFl[k * (NT + 4) + tid];
            if (f == 0xff) continue;
            double gp[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) gp[j] = 0.0;
            if (f & ROW_START) {
                g = adj_zero<ND>();
            } else {
                State<ND> sk;
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    sk.a[d].x = sFs[SM::SF::at(k, 2 * d, tid)];
                    sk.a[d].y = sFs[SM::SF::at(k, 2 * d + 1, tid)];
                }
                sk.P.a = sFs[SM::SF::at(k, 2 * ND, tid)];
                sk.P.b = sFs[SM::SF::at(k, 2 * ND + 1, tid)];
                sk.P.c = sFs[SM::SF::at(k, 2 * ND + 2, tid)];
                double mu[ND], y[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    mu[d] = sW[SM::SW::at(k, d, tid)];
                    y[d] = sY[SM::SY::at(k, d, tid)];
                }
                const double tau = sW[SM::SW::at(k, ND, tid)], e = sW[SM::SW::at(k, ND + 1, tid)],
                             s2 = sW[SM::SW::at(k, ND + 2, tid)], dt = sDt[SM::SD::at(k, 0, tid)];
                const StepPar sp = make_step(tau, e, s2, dt);
                const bool has = (f & ROW_OBS) != 0, cut = (f & ROW_LAST) != 0;
                StepAux<ND> ax;
                fwd_step<ND, true>(sk, sp, y, mu, has, h, &ax);
                const Adj<ND> gin = cut ? adj_zero<ND>() : g;
                double g_h;
                row_param_grad<ND>(gin, sp, ax, mu, tau, e, s2, dt, has, gp, gp[ND], gp[ND + 1], g_h);
                gh += g_h;
                g = bwd_apply<ND>(bwd_row_elem<ND>(sp, ax, has, cut), g);
            }
            double* out = a.eta_bar + (size_t)(row0 + k) * NP;
#pragma unroll
            for (int j = 0; j < NP; ++j) out[j] = gp[j];
        }
        const double tg = block_sum<NT>(gh, sMisc + 16);
        if (tid == 0) a.tile_gh[tile] = tg;
    }
}

}  // namespace ssde
