// Device layout of the data list and of the sparse spline design matrices, and the two design
// products every kernel fuses in: the linear predictor eta_i = X_i theta (replaces
// par_vec = X_fe*coeff_fe + X_re*coeff_re and its reshape, nllk_ctcrw.hpp:143-149,
// nllk_sde.hpp:61-67) and the transposed product grad_theta += X' eta_bar (TMB's reverse sweep of
// the same line).
//
// "Warp-tile transposed" layout, built once in ssde_create:
//   * rows are grouped in warp-tiles of WT = 32 * LC consecutive rows; lane l of the warp that
//     processes a warp-tile owns the LC consecutive rows  q*WT + l*LC + k, k = 0..LC-1 (a
//     contiguous piece of the time series, which is what the scan kernels need);
//   * every per-row array (flags, dt, obs planes) is stored at  pos = q*WT + k*32 + l, so the
//     k-th row of all 32 lanes is one coalesced 256-byte (doubles) request straight into
//     registers -- no shared-memory staging;
//   * the design is stored per warp-tile in a sliced-ELL form: X_fe and X_re are block-diagonal
//     by parameter (rows j*n..(j+1)*n-1 belong to parameter j, R/sde.R:443-447), so row i has a
//     short list of nonzeros per parameter.  A warp-tile has kmax_p slots for parameter p
//     (S = sum_p kmax_p), value slot j of row (k, l) sits at  val[val_off + (k*S + j)*32 + l];
//   * mgcv smooth blocks are dense in their rows and `bs = "re"` blocks are constant along a
//     track, so almost every warp-tile uses the same columns in all of its rows: then
//     (WT_UNIFORM) the column of slot j is stored once per warp-tile, col[col_off + j], instead
//     of once per nonzero (zeros are filled in for rows that lack a column of the union).
//     Otherwise columns are stored per nonzero with the same indexing as val.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace ssde {

constexpr int LC = 8;              // rows per thread chunk
constexpr int WT = 32 * LC;        // rows per warp-tile
constexpr int MAX_NP = 4;          // SDE parameters per row (n_dim + 2 <= 4)
constexpr int TH_CACHE = 64;       // per-warp cache of theta[col[j]] for uniform warp-tiles
constexpr int STAGE_SLOTS = 24;    // slots of one row-step that a warp stages in shared memory by TMA
constexpr int STAGE_DBL = STAGE_SLOTS * 32;

enum : uint32_t { WT_UNIFORM = 1u };

struct WtDesc {
    int64_t val_off;
    int64_t col_off;
    uint32_t kmax;                 // byte p = slots of parameter p
    uint32_t flags;
};
static_assert(sizeof(WtDesc) == 24, "descriptor layout is part of the C ABI (ssde_create_packed)");

struct DesignV2 {
    int64_t n;                     // real rows
    int64_t n_pad;                 // rows rounded up to a multiple of WT
    const WtDesc* desc;            // [n_pad / WT]
    const double* val;
    const uint32_t* col;
};

__host__ __device__ __forceinline__ int64_t row_pos(int64_t row) {
    const int64_t q = row / WT;
    const int r = (int)(row - q * WT);
    return q * WT + (int64_t)(r % LC) * 32 + r / LC;
}

__device__ __forceinline__ int slots_of(uint32_t kmax) {
    return (int)((kmax & 255u) + ((kmax >> 8) & 255u) + ((kmax >> 16) & 255u) + (kmax >> 24));
}

// The coefficient vector theta = [coeff_fe | coeff_re] and, for the tangent pass (R = Dual), the
// direction it is differentiated along.
struct Theta {
    const double* v;
    const double* d;               // nullptr unless R = Dual
};
template <class R>
__device__ __forceinline__ R theta_at(const Theta& t, uint32_t c) {
    if constexpr (std::is_same<R, double>::value) return __ldg(t.v + c);
    else return Dual(__ldg(t.v + c), __ldg(t.d + c));
}

// Per-warp view of one warp-tile's design.
template <class R>
struct WtViewT {
    const double* blk;             // val + val_off
    const double* v;               // val + val_off + lane
    const uint32_t* c;             // col + col_off (+ lane if not uniform)
    const R* th;                   // per-warp theta cache (shared memory) or nullptr
    int th_pad;                    // 16: parameter p's cached thetas start at th[16 p]; 0: contiguous
    uint32_t kmax;
    int S;
    bool uniform;
    bool staged;                   // row-steps are fetched with TMA bulk copies (uniform, 0 < S <= STAGE_SLOTS)
};
using WtView = WtViewT<double>;

// Per-warp staging buffer for the values of one row-step (S x 32 doubles, slot-major) and the
// mbarrier its TMA copies complete on.
struct WarpStage {
    double* buf;
    uint64_t* bar;
    unsigned phase;
};

__device__ __forceinline__ void stage_init(WarpStage& st, double* buf, uint64_t* bar) {
    st.buf = buf; st.bar = bar; st.phase = 0;
    if ((threadIdx.x & 31) == 0) mbar_init(bar, 1);
}

// Called by all lanes of a warp.  `th_cache` is this warp's TH_CACHE doubles of shared memory.
template <class R>
__device__ __forceinline__ WtViewT<R> open_warptile(const DesignV2& X, int64_t q, const Theta& theta,
                                                   R* th_cache, bool want_theta = true) {
    const int lane = threadIdx.x & 31;
    const WtDesc d = X.desc[q];
    WtViewT<R> w;
    w.kmax = d.kmax;
    w.S = slots_of(d.kmax);
    w.uniform = (d.flags & WT_UNIFORM) != 0;
    w.blk = X.val + d.val_off;
    w.v = w.blk + lane;
    w.c = X.col + d.col_off + (w.uniform ? 0 : lane);
    w.th = nullptr;
    w.th_pad = 0;
    // every parameter has at most 16 slots -> 16-aligned per-parameter segments (vector loads)
    bool le16 = true;
#pragma unroll
    for (int p = 0; p < MAX_NP; ++p) le16 = le16 && (((d.kmax >> (8 * p)) & 255u) <= 16u);
    if (!want_theta) {
        // caller only walks the values / columns (transposed product)
    } else if (w.uniform && w.S > 0 && w.S <= STAGE_SLOTS && le16) {
        __syncwarp();
        int j0 = 0;
#pragma unroll
        for (int p = 0; p < MAX_NP; ++p) {
            const int kp = (int)((d.kmax >> (8 * p)) & 255u);
            if (lane < kp) th_cache[16 * p + lane] = theta_at<R>(theta, __ldg(w.c + j0 + lane));
            j0 += kp;
        }
        __syncwarp();
        w.th = th_cache;
        w.th_pad = 16;
    } else if (w.uniform && w.S <= TH_CACHE) {
        __syncwarp();
        for (int j = lane; j < w.S; j += 32) th_cache[j] = theta_at<R>(theta, __ldg(w.c + j));
        __syncwarp();
        w.th = th_cache;
    }
    w.staged = w.th_pad != 0;
    return w;
}

// lane 0: start the copy of row-step k of the warp-tile into the staging buffer.  The caller has
// made sure (with __syncwarp) that no lane still reads the buffer.
template <class R>
__device__ __forceinline__ void stage_issue(const WtViewT<R>& w, const WarpStage& st, int k) {
    const unsigned bytes = (unsigned)w.S * 32u * 8u;
    fence_proxy_async();
    mbar_expect_tx(st.bar, bytes);
    tma_load_1d(st.buf, w.blk + (size_t)k * w.S * 32, bytes, st.bar);
}
__device__ __forceinline__ void stage_wait(WarpStage& st) {
    mbar_wait(st.bar, st.phase);
    st.phase ^= 1u;
}

// eta[p] of this lane's row from the staged row-step
template <int NP>
__device__ __forceinline__ void row_eta_staged(const WtViewT<Dual>& w, const WarpStage& st, Dual* eta) {
    const double* v = st.buf + (threadIdx.x & 31);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int kp = (int)((w.kmax >> (8 * p)) & 255u);
        const Dual* th = w.th + 16 * p;
        Dual acc0 = 0.0, acc1 = 0.0;
        int i = 0;
#pragma unroll 1
        for (; i + 2 <= kp; i += 2) {
            acc0 = fmad(v[0], th[i], acc0);
            acc1 = fmad(v[32], th[i + 1], acc1);
            v += 64;
        }
        if (i < kp) { acc0 = fmad(v[0], th[i], acc0); v += 32; }
        eta[p] = acc0 + acc1;
    }
}
template <int NP>
__device__ __forceinline__ void row_eta_staged(const WtView& w, const WarpStage& st, double* eta) {
    const double* v = st.buf + (threadIdx.x & 31);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int kp = (int)((w.kmax >> (8 * p)) & 255u);
        const double* th = w.th + 16 * p;
        double acc0 = 0.0, acc1 = 0.0;
        int i = 0;
#pragma unroll 1
        for (; i + 4 <= kp; i += 4) {
            const double2 ta = *reinterpret_cast<const double2*>(th + i);
            const double2 tb = *reinterpret_cast<const double2*>(th + i + 2);
            acc0 = fma(v[0], ta.x, acc0);
            acc1 = fma(v[32], ta.y, acc1);
            acc0 = fma(v[64], tb.x, acc0);
            acc1 = fma(v[96], tb.y, acc1);
            v += 128;
        }
        if (kp - i >= 2) {
            const double2 ta = *reinterpret_cast<const double2*>(th + i);
            acc0 = fma(v[0], ta.x, acc0);
            acc1 = fma(v[32], ta.y, acc1);
            v += 64; i += 2;
        }
        if (i < kp) { acc0 = fma(v[0], th[i], acc0); v += 32; }
        eta[p] = acc0 + acc1;
    }
}

// eta[p] for the first NPRE parameters only, straight from global memory (row k of this lane)
template <int NPRE, class R>
__device__ __forceinline__ void row_eta_prefix(const WtViewT<R>& w, int k, const Theta& theta, R* eta) {
    const double* v = w.v + (size_t)k * w.S * 32;
    const uint32_t* c = w.c + (w.uniform ? 0 : (size_t)k * w.S * 32);
    const int cs = w.uniform ? 1 : 32;
    int j = 0;
#pragma unroll
    for (int p = 0; p < NPRE; ++p) {
        const int kp = (int)((w.kmax >> (8 * p)) & 255u);
        R acc = 0.0;
        for (int i = 0; i < kp; ++i, ++j) {
            const R t = w.th ? w.th[w.th_pad ? 16 * p + i : j] : theta_at<R>(theta, __ldg(c + j * cs));
            acc = fmad(__ldg(v + j * 32), t, acc);
        }
        eta[p] = acc;
    }
}

// eta[p] = sum over parameter p's slots of row k of this lane.
template <int NP, class R>
__device__ __forceinline__ void row_eta(const WtViewT<R>& w, int k, const Theta& theta, R* eta) {
    const double* v = w.v + (size_t)k * w.S * 32;
    int j = 0;
    if (w.th) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int kp = (int)((w.kmax >> (8 * p)) & 255u);
            R acc = 0.0;
#pragma unroll 4
            for (int i = 0; i < kp; ++i, ++j) acc = fmad(__ldg(v + j * 32), w.th[w.th_pad ? 16 * p + i : j], acc);
            eta[p] = acc;
        }
    } else {
        const uint32_t* c = w.c + (w.uniform ? 0 : (size_t)k * w.S * 32);
        const int cs = w.uniform ? 1 : 32;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int kp = (int)((w.kmax >> (8 * p)) & 255u);
            R acc = 0.0;
            for (int i = 0; i < kp; ++i, ++j) acc = fmad(__ldg(v + j * 32), theta_at<R>(theta, __ldg(c + j * cs)), acc);
            eta[p] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// transposed product for one warp-tile.  eb(k, p) returns d nllk / d eta of this lane's row k,
// parameter p.  Per-CTA accumulators `sgrad` (shared memory, npc doubles, flushed by the caller)
// are used when the parameter vector fits, global atomics otherwise.
// ---------------------------------------------------------------------------------------------
template <class R>
struct GradAccT {
    R* sgrad;                      // shared-memory accumulators or nullptr
    double* ggrad;                 // global gradient w.r.t. theta; R = Dual: tangents at ggrad[p_theta + c]
    int p_theta;
};
using GradAcc = GradAccT<double>;

__device__ __forceinline__ void grad_add(const GradAcc& g, uint32_t c, double v) {
    if (g.sgrad) atomicAdd(g.sgrad + c, v);
    else atomicAdd(g.ggrad + c, v);
}
__device__ __forceinline__ void grad_add(const GradAccT<Dual>& g, uint32_t c, const Dual& v) {
    if (g.sgrad) { atomicAdd(&g.sgrad[c].v, v.v); atomicAdd(&g.sgrad[c].d, v.d); }
    else { atomicAdd(g.ggrad + c, v.v); atomicAdd(g.ggrad + g.p_theta + c, v.d); }
}
// flush the per-CTA accumulators (all threads of the CTA, after a __syncthreads)
template <class R>
__device__ __forceinline__ void grad_flush(const GradAccT<R>& g, int nthreads) {
    if (!g.sgrad) return;
    for (int i = threadIdx.x; i < g.p_theta; i += nthreads) {
        const R v = g.sgrad[i];
        if (value(v) != 0.0) atomicAdd(g.ggrad + i, value(v));
        if constexpr (!std::is_same<R, double>::value) { if (v.d != 0.0) atomicAdd(g.ggrad + g.p_theta + i, v.d); }
    }
}
__device__ __forceinline__ bool nonzero(double x) { return x != 0.0; }
__device__ __forceinline__ bool nonzero(const Dual& x) { return x.v != 0.0 || x.d != 0.0; }

// Uniform warp-tile: per-lane partial sums over the LC rows go to a per-warp scratch T(j, lane)
// in shared memory (accessor `T`), then lane j adds up row j of T -- no shuffles, one atomic per
// slot.  The caller guarantees that T has room for w.S slots.
#ifndef SSDE_SCATTER_BATCH
#define SSDE_SCATTER_BATCH 4
#endif
template <int NP, class R, class EB, class TA>
__device__ __forceinline__ void scatter_warptile_transposed(const WtViewT<R>& w, const GradAccT<R>& g, EB eb, TA T) {
    constexpr int SB = SSDE_SCATTER_BATCH;
    const int lane = threadIdx.x & 31;
    const size_t ks = (size_t)w.S * 32;
    const double* vj = w.v;
    int j = 0;
    __syncwarp();
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int kp = (int)((w.kmax >> (8 * p)) & 255u);
        R e[LC];
#pragma unroll
        for (int k = 0; k < LC; ++k) e[k] = eb(k, p);
        // SB slots x LC rows independent loads in flight per batch
#pragma unroll 1
        for (int i = 0; i < kp; i += SB) {
            double v[SB][LC];
#pragma unroll
            for (int u = 0; u < SB; ++u)
#pragma unroll
                for (int k = 0; k < LC; ++k) v[u][k] = (i + u < kp) ? __ldg(vj + u * 32 + k * ks) : 0.0;
#pragma unroll
            for (int u = 0; u < SB; ++u) {
                R acc = 0.0;
#pragma unroll
                for (int k = 0; k < LC; ++k) acc = fmad(v[u][k], e[k], acc);
                if (i + u < kp) T(j + u, lane) = acc;
            }
            vj += SB * 32;
            j += SB;
        }
        // slots were advanced in steps of SB: step back to the first slot of the next parameter
        {
            const int over = (SB - kp % SB) % SB;
            vj -= over * 32;
            j -= over;
        }
    }
    __syncwarp();
    for (int jj = lane; jj < w.S; jj += 32) {
        R s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            s0 += T(jj, (i + lane) & 31);
            s1 += T(jj, (i + 1 + lane) & 31);
        }
        grad_add(g, __ldg(w.c + jj), s0 + s1);
    }
    __syncwarp();
}

template <int NP, class R, class EB>
__device__ __forceinline__ void scatter_warptile(const WtViewT<R>& w, const GradAccT<R>& g, EB eb) {
    const int lane = threadIdx.x & 31;
    int j = 0;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int kp = (int)((w.kmax >> (8 * p)) & 255u);
        R e[LC];
#pragma unroll
        for (int k = 0; k < LC; ++k) e[k] = eb(k, p);
        for (int i = 0; i < kp; ++i, ++j) {
            if (w.uniform) {
                R acc = 0.0;
#pragma unroll
                for (int k = 0; k < LC; ++k) acc = fmad(__ldg(w.v + ((size_t)k * w.S + j) * 32), e[k], acc);
                acc = warp_sum(acc);
                if (lane == 0) grad_add(g, __ldg(w.c + j), acc);
            } else {
#pragma unroll
                for (int k = 0; k < LC; ++k) {
                    const size_t o = ((size_t)k * w.S + j) * 32;
                    const R t = __ldg(w.v + o) * e[k];
                    if (nonzero(t)) grad_add(g, __ldg(w.c + o), t);
                }
            }
        }
    }
}

}  // namespace ssde
