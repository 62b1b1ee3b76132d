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
//     Otherwise columns are stored per nonzero with the same indexing as val;
//   * formulas usually give several SDE parameters the SAME smooth (tau ~ s(time), nu ~ s(time)):
//     their blocks of X_re hold identical numbers in different columns.  In a uniform warp-tile a
//     parameter p whose value slots equal parameter t's (t < p) in every row is an ALIAS of t
//     (descriptor flags, alias_of): it keeps its own column slots but stores no values, so a row
//     carries SV = sum of the non-aliased kmax_p value slots (stride of a row-step: SV * 32) while
//     the column list still has S = sum_p kmax_p entries.  Benchmark shape: 22 -> 11 doubles per row.
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

enum : uint32_t { WT_UNIFORM = 1u, WT_ALIAS_SHIFT = 8u, WT_ALIAS_MASK = 0xff00u };
// bits 8+2p, 9+2p of the descriptor flags: 0 = parameter p stores its own values, t + 1 = its value
// slots are those of parameter t (t < p, t not an alias itself, kmax_t == kmax_p, uniform warp-tile)
__host__ __device__ __forceinline__ int alias_of(uint32_t flags, int p) {
    return (int)((flags >> (WT_ALIAS_SHIFT + 2 * p)) & 3u) - 1;
}

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

__host__ __device__ __forceinline__ int slots_of(uint32_t kmax) {
    return (int)((kmax & 255u) + ((kmax >> 8) & 255u) + ((kmax >> 16) & 255u) + (kmax >> 24));
}
// value slots per row (aliased parameters store none); vofs: 16 bits per parameter = first value
// slot of parameter p (an alias points at its target's)
__host__ __device__ __forceinline__ int value_slots_of(uint32_t kmax, uint32_t flags, uint64_t* vofs = nullptr) {
    int sv = 0;
    uint64_t o = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int t = alias_of(flags, p);
        if (t >= 0) {
            o |= ((o >> (16 * t)) & 0xffffull) << (16 * p);
        } else {
            o |= (uint64_t)sv << (16 * p);
            sv += (int)((kmax >> (8 * p)) & 255u);
        }
    }
    if (vofs) *vofs = o;
    return sv;
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
    uint32_t flags;                // descriptor flags (alias bits)
    uint64_t vofs;                 // 16 bits per parameter: its first value slot
    int S;                         // column slots
    int SV;                        // value slots of a row (== S unless parameters are aliased)
    bool uniform;
    bool staged;                   // row-steps are fetched with TMA bulk copies (uniform, 0 < SV <= STAGE_SLOTS)
    __device__ __forceinline__ int voff(int p) const { return (int)((vofs >> (16 * p)) & 0xffffull); }
    // parameter p + 1 is an alias of p: both predictors come out of one pass over p's values
    __device__ __forceinline__ bool pair(int p) const { return p + 1 < MAX_NP && alias_of(flags, p + 1) == p; }
    __device__ __forceinline__ bool second_of_pair(int p) const { return p > 0 && alias_of(flags, p) == p - 1; }
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

// What a warp's theta cache currently holds: consecutive warp-tiles of mgcv-smooth designs share
// one column list (same col_off), and theta does not change during a launch, so the cache filled
// for the previous warp-tile is reused -- this removes two of the three dependent global loads
// (descriptor -> columns -> theta) from the start of a tile.
struct ThetaKey {
    int64_t col_off = -1;
    uint32_t kmax = 0, flags = 0;
    bool fresh = false;            // set by open_warptile: the cache was (re)filled for this warp-tile
};

// Called by all lanes of a warp.  `th_cache` is this warp's TH_CACHE doubles of shared memory.
template <class R>
__device__ __forceinline__ WtViewT<R> open_warptile(const DesignV2& X, int64_t q, const Theta& theta,
                                                   R* th_cache, bool want_theta = true, ThetaKey* key = nullptr) {
    const int lane = threadIdx.x & 31;
    const WtDesc d = X.desc[q];
    WtViewT<R> w;
    w.kmax = d.kmax;
    w.flags = d.flags;
    w.S = slots_of(d.kmax);
    w.SV = value_slots_of(d.kmax, d.flags, &w.vofs);
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
    } else if (w.uniform && w.SV > 0 && w.SV <= STAGE_SLOTS && le16) {
        const bool same = key && key->col_off == d.col_off && key->kmax == d.kmax && key->flags == d.flags;
        if (key) { key->col_off = d.col_off; key->kmax = d.kmax; key->flags = d.flags; key->fresh = !same; }
        if (!same) {
            __syncwarp();
            int j0 = 0;
#pragma unroll
            for (int p = 0; p < MAX_NP; ++p) {
                const int kp = (int)((d.kmax >> (8 * p)) & 255u);
                if (lane < kp) th_cache[16 * p + lane] = theta_at<R>(theta, __ldg(w.c + j0 + lane));
                j0 += kp;
            }
            __syncwarp();
        }
        w.th = th_cache;
        w.th_pad = 16;
    } else if (w.uniform && w.S <= TH_CACHE) {
        __syncwarp();
        for (int j = lane; j < w.S; j += 32) th_cache[j] = theta_at<R>(theta, __ldg(w.c + j));
        __syncwarp();
        w.th = th_cache;
    }
    w.staged = w.th_pad != 0;
    if (key && !w.staged) { key->col_off = -1; key->fresh = true; }
    return w;
}

// Dense coefficient table of a staged warp-tile: TH[j][p] = coefficient of value slot j in parameter
// p's predictor (theta of p's column there, 0 where the slot is not p's; an aliased parameter has
// its coefficients in its target's rows).  With it a row's predictors are ONE branch-free pass over
// the SV staged values, eta[p] += x_j TH[j][p] -- the per-parameter loops with their remainder
// handling cost ~270 instructions per row, this ~100 (adding x * 0 leaves a sum unchanged).
__device__ __forceinline__ void fill_theta_matrix(const WtView& w, double* thm) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    // rows up to the next multiple of four are zero: row_eta_dense walks the slots in groups of four
    for (int i = lane; i < ((w.SV + 3) & ~3) * MAX_NP; i += 32) thm[i] = 0.0;
    __syncwarp();
#pragma unroll
    for (int p = 0; p < MAX_NP; ++p) {
        const int kp = (int)((w.kmax >> (8 * p)) & 255u);
        if (lane < kp) thm[(w.voff(p) + lane) * MAX_NP + p] = w.th[16 * p + lane];
    }
    __syncwarp();
}
// SKIP: the first SKIP parameters are known to have all-zero coefficients (mu fixed at 0, the
// kernels' mu_zero flag): their sums are not formed.  Slots are taken four at a time with constant
// offsets; the up to three slots past SV meet zero table rows, and the staging buffer only ever
// holds finite numbers there (zeroed at kernel start, then design values or the scan elements the
// forward kernel parks in it -- finite unless the filter itself has already overflowed, in which
// case the evaluation is NaN either way).
template <int NP, int SKIP>
__device__ __forceinline__ void row_eta_dense(const WtView& w, const double* buf, const double* thm, double* eta) {
    static_assert(NP <= MAX_NP && MAX_NP == 4 && SKIP < NP, "two double2 per table row");
    const double* v = buf + (threadIdx.x & 31);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const int groups = (w.SV + 3) >> 2;
#pragma unroll 1
    for (int g = 0; g < groups; ++g) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double x = v[u * 32];
            if (SKIP < 2) {
                const double2 t01 = *reinterpret_cast<const double2*>(thm + u * MAX_NP);
                if (SKIP < 1) a0 = fma(x, t01.x, a0);
                if (NP > 1) a1 = fma(x, t01.y, a1);
            }
            if (NP > 2) {
                const double2 t23 = *reinterpret_cast<const double2*>(thm + u * MAX_NP + 2);
                if (SKIP < 3) a2 = fma(x, t23.x, a2);
                if (NP > 3) a3 = fma(x, t23.y, a3);
            }
        }
        v += 4 * 32;
        thm += 4 * MAX_NP;
    }
    eta[0] = a0;
    if (NP > 1) eta[1] = a1;
    if (NP > 2) eta[2] = a2;
    if (NP > 3) eta[3] = a3;
}

// lane 0: start the copy of row-step k of the warp-tile into the staging buffer.  The caller has
// made sure (with __syncwarp) that no lane still reads the buffer.
template <class R>
__device__ __forceinline__ void stage_issue(const WtViewT<R>& w, const WarpStage& st, int k) {
    const unsigned bytes = (unsigned)w.SV * 32u * 8u;
    fence_proxy_async();
    mbar_expect_tx(st.bar, bytes);
    tma_load_1d(st.buf, w.blk + (size_t)k * w.SV * 32, bytes, st.bar);
}
// the same with the source pointer / size kept by the caller (no index arithmetic per row-step)
__device__ __forceinline__ void stage_issue_at(const WarpStage& st, const double* src, unsigned bytes) {
    fence_proxy_async();
    mbar_expect_tx(st.bar, bytes);
    tma_load_1d(st.buf, src, bytes, st.bar);
}
__device__ __forceinline__ void stage_wait(WarpStage& st) {
    mbar_wait(st.bar, st.phase);
    st.phase ^= 1u;
}

// eta[p] of this lane's row from the staged row-step (`buf` = the row-step's SV x 32 values)
template <int NP>
__device__ __forceinline__ void row_eta_staged(const WtViewT<Dual>& w, const double* buf, Dual* eta) {
    const double* v0 = buf + (threadIdx.x & 31);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int kp = (int)((w.kmax >> (8 * p)) & 255u);
        const double* v = v0 + w.voff(p) * 32;
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
__device__ __forceinline__ void row_eta_staged(const WtView& w, const double* buf, double* eta) {
    const double* v0 = buf + (threadIdx.x & 31);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        if (w.second_of_pair(p)) continue;             // came out of parameter p - 1's pass
        const int kp = (int)((w.kmax >> (8 * p)) & 255u);
        const double* v = v0 + w.voff(p) * 32;
        const double* th = w.th + 16 * p;
        // even slots accumulate in acc0, odd slots in acc1 (the same order with and without aliases)
        double acc0 = 0.0, acc1 = 0.0;
        int i = 0;
        if (p + 1 < NP && w.pair(p)) {
            const double* tg = th + 16;
            double bcc0 = 0.0, bcc1 = 0.0;
#pragma unroll 1
            for (; i + 4 <= kp; i += 4) {
                const double2 ta = *reinterpret_cast<const double2*>(th + i);
                const double2 tb = *reinterpret_cast<const double2*>(th + i + 2);
                const double2 ga = *reinterpret_cast<const double2*>(tg + i);
                const double2 gb = *reinterpret_cast<const double2*>(tg + i + 2);
                const double x0 = v[0], x1 = v[32], x2 = v[64], x3 = v[96];
                acc0 = fma(x0, ta.x, acc0); bcc0 = fma(x0, ga.x, bcc0);
                acc1 = fma(x1, ta.y, acc1); bcc1 = fma(x1, ga.y, bcc1);
                acc0 = fma(x2, tb.x, acc0); bcc0 = fma(x2, gb.x, bcc0);
                acc1 = fma(x3, tb.y, acc1); bcc1 = fma(x3, gb.y, bcc1);
                v += 128;
            }
            if (kp - i >= 2) {
                const double2 ta = *reinterpret_cast<const double2*>(th + i);
                const double2 ga = *reinterpret_cast<const double2*>(tg + i);
                const double x0 = v[0], x1 = v[32];
                acc0 = fma(x0, ta.x, acc0); bcc0 = fma(x0, ga.x, bcc0);
                acc1 = fma(x1, ta.y, acc1); bcc1 = fma(x1, ga.y, bcc1);
                v += 64; i += 2;
            }
            if (i < kp) { const double x0 = v[0]; acc0 = fma(x0, th[i], acc0); bcc0 = fma(x0, tg[i], bcc0); }
            eta[p] = acc0 + acc1;
            if (p + 1 < NP) eta[p + 1] = bcc0 + bcc1;
            continue;
        }
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
template <int NP, class R>
__device__ __forceinline__ void row_eta_staged(const WtViewT<R>& w, const WarpStage& st, R* eta) {
    row_eta_staged<NP>(w, st.buf, eta);
}

// eta[p] for the first NPRE parameters only, straight from global memory (row k of this lane)
template <int NPRE, class R>
__device__ __forceinline__ void row_eta_prefix(const WtViewT<R>& w, int k, const Theta& theta, R* eta) {
    const double* vk = w.v + (size_t)k * w.SV * 32;
    const uint32_t* c = w.c + (w.uniform ? 0 : (size_t)k * w.S * 32);
    const int cs = w.uniform ? 1 : 32;
    int j = 0;
#pragma unroll
    for (int p = 0; p < NPRE; ++p) {
        const int kp = (int)((w.kmax >> (8 * p)) & 255u);
        const double* v = vk + w.voff(p) * 32;
        R acc = 0.0;
        for (int i = 0; i < kp; ++i, ++j) {
            const R t = w.th ? w.th[w.th_pad ? 16 * p + i : j] : theta_at<R>(theta, __ldg(c + j * cs));
            acc = fmad(__ldg(v + i * 32), t, acc);
        }
        eta[p] = acc;
    }
}

// eta[p] = sum over parameter p's slots of row k of this lane.
template <int NP, class R>
__device__ __forceinline__ void row_eta(const WtViewT<R>& w, int k, const Theta& theta, R* eta) {
    const double* vk = w.v + (size_t)k * w.SV * 32;
    int j = 0;
    if (w.th) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int kp = (int)((w.kmax >> (8 * p)) & 255u);
            const double* v = vk + w.voff(p) * 32;
            R acc = 0.0;
#pragma unroll 4
            for (int i = 0; i < kp; ++i, ++j) acc = fmad(__ldg(v + i * 32), w.th[w.th_pad ? 16 * p + i : j], acc);
            eta[p] = acc;
        }
    } else {
        const uint32_t* c = w.c + (w.uniform ? 0 : (size_t)k * w.S * 32);
        const int cs = w.uniform ? 1 : 32;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int kp = (int)((w.kmax >> (8 * p)) & 255u);
            const double* v = vk + w.voff(p) * 32;
            R acc = 0.0;
            for (int i = 0; i < kp; ++i, ++j) acc = fmad(__ldg(v + i * 32), theta_at<R>(theta, __ldg(c + j * cs)), acc);
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
#ifndef SSDE_SCATTER_BATCH2
#define SSDE_SCATTER_BATCH2 4          // slots per batch when two parameters share the loaded values
#endif
template <int NP, class R, class EB, class TA>
__device__ __forceinline__ void scatter_warptile_transposed(const WtViewT<R>& w, const GradAccT<R>& g, EB eb, TA T) {
    constexpr int SB = SSDE_SCATTER_BATCH;
    const int lane = threadIdx.x & 31;
    const size_t ks = (size_t)w.SV * 32;
    int jc = 0;                                    // first column slot of parameter p
    __syncwarp();
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int kp = (int)((w.kmax >> (8 * p)) & 255u);
        if (w.second_of_pair(p)) { jc += kp; continue; }      // done together with parameter p - 1
        const double* vj = w.v + w.voff(p) * 32;
        R e[LC];
#pragma unroll
        for (int k = 0; k < LC; ++k) e[k] = eb(k, p);
        if (p + 1 < NP && w.pair(p)) {
            // parameter p + 1 uses the same values: one load feeds both products
            constexpr int SB2 = SSDE_SCATTER_BATCH2;
            R e2[LC];
#pragma unroll
            for (int k = 0; k < LC; ++k) e2[k] = eb(k, p + 1 < NP ? p + 1 : p);
#pragma unroll 1
            for (int i = 0; i < kp; i += SB2) {
                double v[SB2][LC];
#pragma unroll
                for (int u = 0; u < SB2; ++u)
#pragma unroll
                    for (int k = 0; k < LC; ++k) v[u][k] = (i + u < kp) ? __ldg(vj + u * 32 + k * ks) : 0.0;
#pragma unroll
                for (int u = 0; u < SB2; ++u) {
                    R acc = 0.0, acc2 = 0.0;
#pragma unroll
                    for (int k = 0; k < LC; ++k) { acc = fmad(v[u][k], e[k], acc); acc2 = fmad(v[u][k], e2[k], acc2); }
                    if (i + u < kp) { T(jc + i + u, lane) = acc; T(jc + kp + i + u, lane) = acc2; }
                }
                vj += SB2 * 32;
            }
            jc += kp;
            continue;
        }
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
                if (i + u < kp) T(jc + i + u, lane) = acc;
            }
            vj += SB * 32;
        }
        jc += kp;
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
        const int jv = w.voff(p);                  // == j unless parameters are aliased (uniform warp-tiles only)
        R e[LC];
#pragma unroll
        for (int k = 0; k < LC; ++k) e[k] = eb(k, p);
        for (int i = 0; i < kp; ++i, ++j) {
            if (w.uniform) {
                R acc = 0.0;
#pragma unroll
                for (int k = 0; k < LC; ++k) acc = fmad(__ldg(w.v + ((size_t)k * w.SV + jv + i) * 32), e[k], acc);
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
