// Fused BM / OU kernel: eta_i = X_i theta -> transition log-density of row i -> row i+1
// (parameters of row i, dt_i = t_{i+1} - t_i; nllk_sde.hpp:77-84 with tr_dens.hpp:32-37 for BM
// and :45-52 for OU) -> closed-form d nllk / d eta -> X' eta_bar, one pass over the data.
// The log-densities are reduced with warp shuffles and one block reduction per CTA.
#pragma once

#include "design.cuh"

namespace ssde {

constexpr int SDE_NT = 128;
constexpr double LOG_SQRT_2PI = 0.918938533204672741780329736406;
enum { MODEL_BM = 0, MODEL_OU = 1, MODEL_CTCRW = 2 };
constexpr int SDE_SGRAD = 512;

struct SdeArgs {
    DesignV2 X;
    Theta theta;               // [coeff_fe | coeff_re] (+ direction for the tangent pass)
    const double* obs;         // ND planes of n_pad doubles, permuted, NA replaced by 0
    const double* dt;          // [n_pad] permuted
    const uint8_t* flags;      // [n_pad] permuted, 0xff beyond the end
    int want_grad;
    double* grad_theta;        // [p_theta] (R = Dual: [2 p_theta], tangents second)
    int p_theta;
    double* block_llk;         // [gridDim.x]
    int64_t ntiles;            // tiles of SDE_NT/32 warp-tiles
};

template <int NP, class R = double>
struct SdeSmem {
    R eb[LC][NP][SDE_NT];
    R th[SDE_NT / 32][TH_CACHE];
    R sgrad[SDE_SGRAD];
    double red[8];
};

// Transition log-density of one row (observation i -> i+1 with the parameters of row i) and its
// derivatives w.r.t. the row's linear predictors: llk += sum over dimensions with both endpoints
// observed (tr_dens.hpp:31), eb[p] += d nllk / d eta_p.  `na` bit d: dimension d is NA at either
// endpoint; z(d, 0 / 1) = observation of dimension d at row i / i+1.
template <int MODEL, int ND, class R, class ZF>
__device__ __forceinline__ void sde_row(const R* eta, double d_t, unsigned na, ZF z, R& llk, R* eb) {
    if (MODEL == MODEL_BM) {
        // mean = z0 + mu dt, sd = exp(eta_s) sqrt(dt)   (tr_dens.hpp:35-36)
        const R sd = exp(eta[ND]) * sqrt(d_t);
        const R isd = 1.0 / sd, lsd = log(sd);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            if ((na >> d) & 1) continue;      // tr_dens.hpp:31
            const double z0 = z(d, 0), z1 = z(d, 1);
            const R res = (z1 - (z0 + eta[d] * d_t)) * isd;
            llk += -LOG_SQRT_2PI - lsd - 0.5 * res * res;
            eb[d] = -res * d_t * isd;
            eb[ND] += 1.0 - res * res;
        }
    } else {
        // mean = mu + exp(-dt/tau)(z0 - mu), sd = sqrt(kappa (1 - exp(-2 dt/tau)))
        const R tau = exp(eta[ND]), kappa = exp(eta[ND + 1]);
        const R ph = exp(-d_t / tau);
        // tr_dens.hpp:50-51 evaluates exp(-2 dt / tau) on its own; ph * ph differs from it by <= 1.5 ulp,
        // far inside the 1e-10 budget, and saves a fourth exp per row
        const R var = kappa * (1.0 - ph * ph);
        const R sd = sqrt(var), isd = 1.0 / sd, lsd = log(sd);
        const R dph = ph * d_t / tau;                  // d ph / d eta_tau
        const R dlv = -2.0 * kappa * ph * dph / var;   // d log var / d eta_tau
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            if ((na >> d) & 1) continue;
            const double z0 = z(d, 0), z1 = z(d, 1);
            const R res = (z1 - (eta[d] + ph * (z0 - eta[d]))) * isd;
            llk += -LOG_SQRT_2PI - lsd - 0.5 * res * res;
            eb[d] = -res * (1.0 - ph) * isd;
            eb[ND] += 0.5 * dlv * (1.0 - res * res) - res * isd * dph * (z0 - eta[d]);
            eb[ND + 1] += 0.5 * (1.0 - res * res);
        }
    }
}

template <int MODEL, int ND, class R = double>
__global__ void __launch_bounds__(SDE_NT) sde_fused_kernel(SdeArgs a) {
    constexpr int NP = (MODEL == MODEL_BM) ? ND + 1 : ND + 2;
    constexpr int NWARP = SDE_NT / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SdeSmem<NP, R>& sm = *reinterpret_cast<SdeSmem<NP, R>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    GradAccT<R> gacc{(a.p_theta <= SDE_SGRAD) ? sm.sgrad : nullptr, a.grad_theta, a.p_theta};
    if (a.want_grad && gacc.sgrad) for (int i = tid; i < SDE_SGRAD; i += SDE_NT) sm.sgrad[i] = 0.0;
    __syncthreads();
    R llk = 0.0;
    for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int64_t q = tile * NWARP + warp;
        const int64_t base = q * WT + lane;
        const int64_t row0 = q * WT + (int64_t)lane * LC;
        const WtViewT<R> w = open_warptile<R>(a.X, q, a.theta, sm.th[warp]);
#pragma unroll 1
        for (int k = 0; k < LC; ++k) {
            const int64_t pos = base + k * 32;
            const uint8_t f0 = a.flags[pos];
            R eb[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) eb[p] = 0.0;
            if (f0 != 0xff && !(f0 & ROW_LAST)) {               // ID(i) == ID(i+1), nllk_sde.hpp:79
                const int64_t npos = (k < LC - 1) ? pos + 32 : row_pos(row0 + k + 1);
                const uint8_t f1 = a.flags[npos];
                R eta[NP];
                row_eta<NP>(w, k, a.theta, eta);
                const double d_t = a.dt[pos];
                sde_row<MODEL, ND>(eta, d_t, (unsigned)((f0 | f1) >> 3),
                                   [&](int d, int nx) { return a.obs[(size_t)d * a.X.n_pad + (nx ? npos : pos)]; }, llk, eb);
            }
            if (a.want_grad) {
#pragma unroll
                for (int p = 0; p < NP; ++p) sm.eb[k][p][tid] = eb[p];
            }
        }
        if (a.want_grad) scatter_warptile<NP, R>(w, gacc, [&](int k, int p) { return sm.eb[k][p][tid]; });
    }
    if (a.want_grad && gacc.sgrad) {
        __syncthreads();
        grad_flush(gacc, SDE_NT);
    }
    const double bl = block_sum<SDE_NT>(value(llk), sm.red);
    if (threadIdx.x == 0) a.block_llk[blockIdx.x] = bl;
}

// ---------------------------------------------------------------------------------------------
// Streaming variant for designs whose warp-tiles are all uniform with at most SDE_SMAX slots (every
// mgcv smooth / random-intercept design of the benchmark shapes).  One pass over HBM:
//   * the S x 32 values of a row-step arrive by TMA bulk copy into a double-buffered per-warp
//     staging area (the copy of row-step k+1 is in flight while the warp works on k), so a warp
//     keeps 6-12 KB of the design stream in flight instead of a handful of 8-byte loads;
//   * eta is formed from shared memory, the transition density and d nllk / d eta in registers, and
//     the row's contribution to X' eta_bar is accumulated at once, per slot, in registers
//     (acc[j] += x_j * eta_bar_p(j)) -- the design is never read a second time;
//   * after the LC rows the 32 lanes' partial sums are added through the (now idle) staging
//     buffer, one lane per slot, and go to per-CTA shared-memory accumulators when the column
//     lies in a "hot" range (fixed effects and small smooth blocks, which every warp-tile
//     touches) or straight to global memory (random intercepts: one column per track).
// ---------------------------------------------------------------------------------------------
constexpr int SDE_SMAX = 32;          // slots per warp-tile the streaming kernel stages
constexpr int SDE_KPM = 12;           // ... and slots per SDE parameter (static register accumulators)
constexpr int SDE_HOT = 512;          // per-CTA hot accumulators (doubles)
constexpr int SDE_MAX_HOT_RANGES = 8;

struct HotRanges {
    int n;
    int lo[SDE_MAX_HOT_RANGES], hi[SDE_MAX_HOT_RANGES], off[SDE_MAX_HOT_RANGES];
};

struct SdeStreamSmem {
    double stage[SDE_NT / 32][2][SDE_SMAX * 32];
    double th[SDE_NT / 32][SDE_SMAX];
    double hot[SDE_HOT];
    uint64_t bar[SDE_NT / 32][2];
    double red[8];
};

// slots of one parameter in groups of four: full groups run without predicates
#define SSDE_SLOT_GROUPS(KP, BODY)                                           \
    _Pragma("unroll") for (int g_ = 0; g_ < SDE_KPM; g_ += 4) {              \
        if (g_ + 4 <= (KP)) {                                                \
            _Pragma("unroll") for (int i = g_; i < g_ + 4; ++i) { BODY; }    \
        } else if (g_ < (KP)) {                                              \
            _Pragma("unroll") for (int i = g_; i < g_ + 4; ++i) if (i < (KP)) { BODY; } \
        }                                                                    \
    }

// The row arithmetic of sde_row for plain doubles with the divisions folded away (every rewrite is
// an identity up to 1-2 ulp: 1/tau = exp(-eta_tau), 1/sd = rsqrt(var), log sd = log(var)/2,
// d log var / d eta_tau = -2 kappa ph dph / var).  Straight-line code, so that the two rows the
// streaming kernel works on at a time interleave in the instruction stream.
template <int MODEL, int ND>
__device__ __forceinline__ void sde_row_fast(const double* eta, double d_t, unsigned na, const double* z0, const double* z1,
                                             double& llk, double* eb) {
    if (MODEL == MODEL_BM) {
        const double isd = exp(-eta[ND]) * rsqrt(d_t);             // sd = exp(eta_s) sqrt(dt), tr_dens.hpp:36
        const double lsd = eta[ND] + 0.5 * log(d_t);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            if ((na >> d) & 1) continue;                           // tr_dens.hpp:31
            const double res = (z1[d] - (z0[d] + eta[d] * d_t)) * isd;
            llk += -LOG_SQRT_2PI - lsd - 0.5 * res * res;
            eb[d] = -res * d_t * isd;
            eb[ND] += 1.0 - res * res;
        }
    } else {
        const double itau = exp(-eta[ND]), kappa = exp(eta[ND + 1]);
        const double x = d_t * itau;
        const double ph = exp(-x);                                 // tr_dens.hpp:49
        const double var = kappa * (1.0 - ph * ph);                // tr_dens.hpp:50-51 (exp(-2 dt/tau) = ph^2 up to 1.5 ulp)
        const double isd = rsqrt(var), lsd = 0.5 * log(var);
        const double dph = ph * x;                                 // d ph / d eta_tau
        const double dlv = -2.0 * kappa * ph * dph * isd * isd;    // d log var / d eta_tau
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            if ((na >> d) & 1) continue;
            const double res = (z1[d] - (eta[d] + ph * (z0[d] - eta[d]))) * isd;
            llk += -LOG_SQRT_2PI - lsd - 0.5 * res * res;
            eb[d] = -res * (1.0 - ph) * isd;
            eb[ND] += 0.5 * dlv * (1.0 - res * res) - res * isd * dph * (z0[d] - eta[d]);
            eb[ND + 1] += 0.5 * (1.0 - res * res);
        }
    }
}

template <int MODEL, int ND>
__global__ void __launch_bounds__(SDE_NT, 3) sde_stream_kernel(SdeArgs a, HotRanges hr) {
    constexpr int NP = (MODEL == MODEL_BM) ? ND + 1 : ND + 2;
    constexpr int NWARP = SDE_NT / 32;
    static_assert(LC % 2 == 0, "rows are processed in pairs");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SdeStreamSmem& sm = *reinterpret_cast<SdeStreamSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < SDE_HOT; i += SDE_NT) sm.hot[i] = 0.0;
    double* buf[2] = {sm.stage[warp][0], sm.stage[warp][1]};
    uint64_t* bar[2] = {&sm.bar[warp][0], &sm.bar[warp][1]};
    if (lane == 0) { mbar_init(bar[0], 1); mbar_init(bar[1], 1); }
    mbar_fence_init();
    __syncthreads();
    unsigned phase = 0;
    double* th = sm.th[warp];
    double llk = 0.0;
    for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int64_t q = tile * NWARP + warp;
        const WtDesc d = a.X.desc[q];
        const int S = slots_of(d.kmax);                            // column slots
        if (S == 0) continue;                                      // padding warp-tile
        uint64_t vofs;
        const int SV = value_slots_of(d.kmax, d.flags, &vofs);     // value slots of a row (< S when parameters share a smooth)
        int kp[NP], vo[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) { kp[p] = (int)((d.kmax >> (8 * p)) & 255u); vo[p] = (int)((vofs >> (16 * p)) & 0xffffull) * 32; }
        const double* blk = a.X.val + d.val_off;
        const uint32_t* cols = a.X.col + d.col_off;
        const unsigned bytes = (unsigned)SV * 32u * 8u;            // one row-step
        __syncwarp();                                              // previous warp-tile: all lanes done with th / staging
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar[0], bytes);
            tma_load_1d(buf[0], blk, bytes, bar[0]);
            mbar_expect_tx(bar[1], bytes);
            tma_load_1d(buf[1], blk + (size_t)SV * 32, bytes, bar[1]);
            // the design stream runs two pairs ahead of the copies: they then hit L2
            prefetch_l2(blk + (size_t)2 * SV * 32, 4u * bytes);
        }
        if (lane < S) th[lane] = __ldg(a.theta.v + __ldg(cols + lane));
        const int64_t base = q * WT + lane;
        const int64_t row0 = q * WT + (int64_t)lane * LC;
        double acc[NP][SDE_KPM];
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
            for (int i = 0; i < SDE_KPM; ++i) acc[p][i] = 0.0;
        // Per-row data: the LC flags of the chunk and the flag of the row after it are loaded up
        // front; dt and the observations of a pair are fetched while the previous pair is worked on.
        const double* p_dt = a.dt + base;
        const double* p_obs[ND];
#pragma unroll
        for (int dd = 0; dd < ND; ++dd) p_obs[dd] = a.obs + (size_t)dd * a.X.n_pad + base;
        const int64_t nx8 = row_pos(row0 + LC) - base;             // offset of the row after this chunk (next lane / next warp-tile)
        const unsigned long long fl = load_flags8(a.flags, base);
        const bool any_next = (uint8_t)(fl >> 56) != 0xff && !((uint8_t)(fl >> 56) & ROW_LAST);
        const uint8_t f8 = any_next ? a.flags[base + nx8] : (uint8_t)0;
        auto live_row = [&](int k) { return (uint8_t)(fl >> (8 * k)) != 0xff; };
        // obs of rows k, k+1, k+2 (za, zb, zc) and dt of rows k, k+1 for the current pair; *_n: next pair
        double za[ND], zb[ND], zc[ND], dta = live_row(0) ? p_dt[0] : 1.0, dtb = live_row(1) ? p_dt[32] : 1.0;
#pragma unroll
        for (int dd = 0; dd < ND; ++dd) {
            za[dd] = live_row(0) ? p_obs[dd][0] : 0.0;
            zb[dd] = live_row(1) ? p_obs[dd][32] : 0.0;
            zc[dd] = live_row(2) ? p_obs[dd][64] : 0.0;
        }
        __syncwarp();
#pragma unroll 1
        for (int k = 0; k < LC; k += 2) {
            // fetch for the next pair (rows k+2, k+3): dt of both, obs of rows k+3 and k+4
            double dta_n = 1.0, dtb_n = 1.0, zb_n[ND], zc_n[ND];
#pragma unroll
            for (int dd = 0; dd < ND; ++dd) { zb_n[dd] = 0.0; zc_n[dd] = 0.0; }
            if (k + 2 < LC) {
                if (live_row(k + 2)) dta_n = p_dt[(k + 2) * 32];
                if (live_row(k + 3)) {
                    dtb_n = p_dt[(k + 3) * 32];
#pragma unroll
                    for (int dd = 0; dd < ND; ++dd) zb_n[dd] = p_obs[dd][(k + 3) * 32];
                }
                if (k + 4 < LC) {
                    if (live_row(k + 4)) {
#pragma unroll
                        for (int dd = 0; dd < ND; ++dd) zc_n[dd] = p_obs[dd][(k + 4) * 32];
                    }
                } else if (any_next) {
#pragma unroll
                    for (int dd = 0; dd < ND; ++dd) zc_n[dd] = p_obs[dd][nx8];
                }
                if (lane == 0 && k + 6 < LC) prefetch_l2(blk + (size_t)(k + 6) * SV * 32, 2u * bytes);
            }
            const uint8_t fa = (uint8_t)(fl >> (8 * k)), fb = (uint8_t)(fl >> (8 * (k + 1)));
            const uint8_t fc = (k + 2 < LC) ? (uint8_t)(fl >> (8 * (k + 2))) : f8;
            const bool live_a = fa != 0xff && !(fa & ROW_LAST);    // ID(i) == ID(i+1), nllk_sde.hpp:79
            const bool live_b = fb != 0xff && !(fb & ROW_LAST);
            mbar_wait(bar[0], phase);
            mbar_wait(bar[1], phase);
            phase ^= 1u;
            const double* va = buf[0] + lane;
            const double* vb = buf[1] + lane;
            double eta_a[NP], eta_b[NP], eb_a[NP], eb_b[NP];
            {
                int o = 0;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const double* xa = va + vo[p];
                    const double* xb = vb + vo[p];
                    const double* tp = th + o;
                    double e0 = 0.0, e1 = 0.0, g0 = 0.0, g1 = 0.0;
                    SSDE_SLOT_GROUPS(kp[p], { const double t = tp[i];
                                              if (i & 1) { e1 = fma(xa[i * 32], t, e1); g1 = fma(xb[i * 32], t, g1); }
                                              else { e0 = fma(xa[i * 32], t, e0); g0 = fma(xb[i * 32], t, g0); } })
                    eta_a[p] = e0 + e1; eta_b[p] = g0 + g1;
                    eb_a[p] = 0.0; eb_b[p] = 0.0;
                    o += kp[p];
                }
            }
            // both rows unconditionally (a dead row has dt = 1 and zero observations: finite arithmetic,
            // result discarded) so that the two dependency chains interleave
            double la = 0.0, lb = 0.0;
            sde_row_fast<MODEL, ND>(eta_a, dta, (unsigned)((fa | fb) >> 3), za, zb, la, eb_a);
            sde_row_fast<MODEL, ND>(eta_b, dtb, (unsigned)((fb | fc) >> 3), zb, zc, lb, eb_b);
            if (live_a) llk += la;
            if (live_b) llk += lb;
            if (a.want_grad) {
                int o = 0;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const double ea = live_a ? eb_a[p] : 0.0, eb_ = live_b ? eb_b[p] : 0.0;
                    const double* xa = va + vo[p];
                    const double* xb = vb + vo[p];
                    SSDE_SLOT_GROUPS(kp[p], acc[p][i] = fma(xb[i * 32], eb_, fma(xa[i * 32], ea, acc[p][i])))
                    o += kp[p];
                }
            }
            __syncwarp();
            if (k + 2 < LC) {
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(bar[0], bytes);
                    tma_load_1d(buf[0], blk + (size_t)(k + 2) * SV * 32, bytes, bar[0]);
                    mbar_expect_tx(bar[1], bytes);
                    tma_load_1d(buf[1], blk + (size_t)(k + 3) * SV * 32, bytes, bar[1]);
                }
                dta = dta_n; dtb = dtb_n;
#pragma unroll
                for (int dd = 0; dd < ND; ++dd) { za[dd] = zc[dd]; zb[dd] = zb_n[dd]; zc[dd] = zc_n[dd]; }
            }
        }
        if (a.want_grad) {
            // X' eta_bar of this warp-tile: T(j, lane) in the staging buffer, lane j adds slot j
            double* T = buf[0] + lane;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                SSDE_SLOT_GROUPS(kp[p], T[i * 32] = acc[p][i])
                T += kp[p] * 32;
            }
            __syncwarp();
            if (lane < S) {
                const double* Tr = buf[0] + lane * 32;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    s0 += Tr[(i + lane) & 31];
                    s1 += Tr[(i + 1 + lane) & 31];
                    s2 += Tr[(i + 2 + lane) & 31];
                    s3 += Tr[(i + 3 + lane) & 31];
                }
                const double sum = (s0 + s1) + (s2 + s3);
                const int c = (int)__ldg(cols + lane);
                int h = -1;
#pragma unroll
                for (int r = 0; r < SDE_MAX_HOT_RANGES; ++r) if (r < hr.n && c >= hr.lo[r] && c < hr.hi[r]) h = hr.off[r] + (c - hr.lo[r]);
                if (h >= 0) atomicAdd(sm.hot + h, sum);
                else atomicAdd(a.grad_theta + c, sum);
            }
        }
    }
    __syncthreads();
    if (a.want_grad) {
        for (int r = 0; r < hr.n; ++r)
            for (int i = tid; i < hr.hi[r] - hr.lo[r]; i += SDE_NT) {
                const double x = sm.hot[hr.off[r] + i];
                if (x != 0.0) atomicAdd(a.grad_theta + hr.lo[r] + i, x);
            }
    }
    const double bl = block_sum<SDE_NT>(llk, sm.red);
    if (threadIdx.x == 0) a.block_llk[blockIdx.x] = bl;
}
#undef SSDE_SLOT_GROUPS

// ---------------------------------------------------------------------------------------------
// Data-term Hessian of the BM / OU models in ONE pass:  H = X' W X  with the exact NP x NP block
// W_i = d^2 nllk_i / d eta_i d eta_i' of every row (rows are independent, nllk_sde.hpp:73-84), the
// "X_re' W X_re" of the Laplace inner problem.  TMB builds the same matrix with its AD-of-AD sparse
// Hessian (MakeADHessObject2, src/init.c:13); tangent passes need one sweep per column, i.e.
// thousands for a random intercept per track (s(ID, bs = "re"), R/sde.R:412-421).
//   * row-step staged by TMA as in sde_stream_kernel; eta per lane (= row) from shared memory;
//   * W_i: the closed-form d nllk_i / d eta of sde_row evaluated on DualN<NP> numbers seeded with
//     the unit directions of the row's predictors -- exact second derivatives, no hand algebra;
//   * the row-step's values are transposed to xT[row][slot] and lane j' accumulates COLUMN j' of
//     the warp-tile's S x S Hessian: acc[P][i] += xT[r][off_P + i] * (W_r[P][p(j')] * xT[r][j'])
//     (broadcast reads of x_j, conflict-free reads of x_j', accumulators statically indexed);
//   * after the LC row-steps column j' is added to per-CTA shared accumulators when both columns are
//     hot (fixed effects, small smooth blocks) and to the dense global matrix otherwise (entries that
//     involve a random intercept: touched by one track's warp-tiles only).
// ---------------------------------------------------------------------------------------------
constexpr int SDE_HESS_HOT = 48;        // hot columns with per-CTA accumulators (HOT x HOT doubles)

struct HessHot {
    int n;                               // hot columns
    int cols[SDE_HESS_HOT];              // their theta indices
};

template <int NP>
struct SdeHessSmem {
    double stage[SDE_NT / 32][SDE_SMAX * 32];
    double xT[SDE_NT / 32][32][SDE_SMAX + 1];
    double W[SDE_NT / 32][32][NP * NP];
    double th[SDE_NT / 32][SDE_SMAX];
    int hidx[SDE_NT / 32][SDE_SMAX];
    double hot[SDE_HESS_HOT * SDE_HESS_HOT];
    uint64_t bar[SDE_NT / 32];
};

#define SSDE_SLOT_GROUPS(KP, BODY)                                           \
    _Pragma("unroll") for (int g_ = 0; g_ < SDE_KPM; g_ += 4) {              \
        if (g_ + 4 <= (KP)) {                                                \
            _Pragma("unroll") for (int i = g_; i < g_ + 4; ++i) { BODY; }    \
        } else if (g_ < (KP)) {                                              \
            _Pragma("unroll") for (int i = g_; i < g_ + 4; ++i) if (i < (KP)) { BODY; } \
        }                                                                    \
    }

template <int MODEL, int ND>
__global__ void __launch_bounds__(SDE_NT, 2) sde_hess_kernel(SdeArgs a, HessHot hh, double* __restrict__ hess) {
    constexpr int NP = (MODEL == MODEL_BM) ? ND + 1 : ND + 2;
    constexpr int NWARP = SDE_NT / 32;
    using DN = DualN<NP>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SdeHessSmem<NP>& sm = *reinterpret_cast<SdeHessSmem<NP>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < SDE_HESS_HOT * SDE_HESS_HOT; i += SDE_NT) sm.hot[i] = 0.0;
    double* stage = sm.stage[warp];
    uint64_t* bar = &sm.bar[warp];
    if (lane == 0) mbar_init(bar, 1);
    mbar_fence_init();
    __syncthreads();
    unsigned phase = 0;
    double* th = sm.th[warp];
    int* hidx = sm.hidx[warp];
    const int64_t p = a.p_theta;
    for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int64_t q = tile * NWARP + warp;
        const WtDesc d = a.X.desc[q];
        const int S = slots_of(d.kmax);                            // column slots
        if (S == 0) continue;
        uint64_t vofs;
        const int SV = value_slots_of(d.kmax, d.flags, &vofs);     // value slots of a row (aliased parameters store none)
        int kp[NP], off[NP], vo[NP];
        {
            int o = 0;
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) {
                kp[pp] = (int)((d.kmax >> (8 * pp)) & 255u); off[pp] = o; o += kp[pp];
                vo[pp] = (int)((vofs >> (16 * pp)) & 0xffffull);
            }
        }
        const double* blk = a.X.val + d.val_off;
        const uint32_t* cols = a.X.col + d.col_off;
        const unsigned bytes = (unsigned)SV * 32u * 8u;
        __syncwarp();
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, bytes);
            tma_load_1d(stage, blk, bytes, bar);
            prefetch_l2(blk + (size_t)SV * 32, (unsigned)((LC - 1) * SV * 32 * 8));
        }
        int my_col = 0, my_q = 0;                                   // this lane's column j' = lane: theta index, SDE parameter
        if (lane < S) {
            my_col = (int)__ldg(cols + lane);
            th[lane] = __ldg(a.theta.v + my_col);
            int h = -1;
            for (int r = 0; r < hh.n; ++r) if (hh.cols[r] == my_col) h = r;
            hidx[lane] = h;
#pragma unroll
            for (int pp = 1; pp < NP; ++pp) my_q += (lane >= off[pp]) ? 1 : 0;
        }
        const int64_t base = q * WT + lane;
        const int64_t row0 = q * WT + (int64_t)lane * LC;
        double acc[NP][SDE_KPM];
#pragma unroll
        for (int pp = 0; pp < NP; ++pp)
#pragma unroll
            for (int i = 0; i < SDE_KPM; ++i) acc[pp][i] = 0.0;
        __syncwarp();
#pragma unroll 1
        for (int k = 0; k < LC; ++k) {
            // this lane's row of the row-step
            const int64_t pos = base + k * 32;
            const uint8_t f0 = a.flags[pos];
            const bool live = f0 != 0xff && !(f0 & ROW_LAST);      // ID(i) == ID(i+1), nllk_sde.hpp:79
            const int64_t npos = (k < LC - 1) ? pos + 32 : row_pos(row0 + k + 1);
            uint8_t f1 = 0;
            double d_t = 1.0, z0[ND], z1[ND];
#pragma unroll
            for (int dd = 0; dd < ND; ++dd) { z0[dd] = 0.0; z1[dd] = 0.0; }
            if (live) {
                f1 = a.flags[npos];
                d_t = a.dt[pos];
#pragma unroll
                for (int dd = 0; dd < ND; ++dd) {
                    z0[dd] = a.obs[(size_t)dd * a.X.n_pad + pos];
                    z1[dd] = a.obs[(size_t)dd * a.X.n_pad + npos];
                }
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            const double* v = stage + lane;
            double eta[NP];
            {
                const double* tp = th;
#pragma unroll
                for (int pp = 0; pp < NP; ++pp) {
                    const double* vp = v + vo[pp] * 32;
                    double e0 = 0.0, e1 = 0.0;
                    SSDE_SLOT_GROUPS(kp[pp], if (i & 1) e1 = fma(vp[i * 32], tp[i], e1); else e0 = fma(vp[i * 32], tp[i], e0))
                    eta[pp] = e0 + e1;
                    tp += kp[pp];
                }
            }
            // transpose the row-step: xT[row = lane][column slot] (an aliased parameter's slots are copies of its target's)
#pragma unroll
            for (int pp = 0; pp < NP; ++pp)
                for (int i = 0; i < kp[pp]; ++i) sm.xT[warp][lane][off[pp] + i] = v[(vo[pp] + i) * 32];
            __syncwarp();
            if (lane == 0 && k + 1 < LC) {                         // the staging buffer is free again
                fence_proxy_async();
                mbar_expect_tx(bar, bytes);
                tma_load_1d(stage, blk + (size_t)(k + 1) * SV * 32, bytes, bar);
            }
            // exact second derivatives of the row's nllk with respect to its predictors
            {
                DN e[NP], eb[NP], llk(0.0);
#pragma unroll
                for (int pp = 0; pp < NP; ++pp) { e[pp] = DN(eta[pp]); e[pp].d[pp] = 1.0; eb[pp] = DN(0.0); }
                if (live) sde_row<MODEL, ND, DN>(e, d_t, (unsigned)((f0 | f1) >> 3), [&](int dd, int nx) { return nx ? z1[dd] : z0[dd]; }, llk, eb);
#pragma unroll
                for (int pp = 0; pp < NP; ++pp)
#pragma unroll
                    for (int qq = 0; qq < NP; ++qq) sm.W[warp][lane][pp * NP + qq] = live ? eb[pp].d[qq] : 0.0;
            }
            __syncwarp();
            // column j' = lane of the warp-tile's X' W X
            if (lane < S) {
#pragma unroll 2
                for (int r = 0; r < 32; ++r) {
                    const double* xr = sm.xT[warp][r];
                    const double* wr = sm.W[warp][r];
                    const double xq = xr[lane];
#pragma unroll
                    for (int pp = 0; pp < NP; ++pp) {
                        const double wx = wr[pp * NP + my_q] * xq;
                        const double* xp = xr + off[pp];
                        SSDE_SLOT_GROUPS(kp[pp], acc[pp][i] = fma(xp[i], wx, acc[pp][i]))
                    }
                }
            }
            __syncwarp();
        }
        // add column j' to the accumulators
        if (lane < S) {
            const int hq = hidx[lane];
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) {
                SSDE_SLOT_GROUPS(kp[pp], {
                    const int j = off[pp] + i;
                    const int hj = hidx[j];
                    const double x = acc[pp][i];
                    if (x != 0.0) {
                        if (hj >= 0 && hq >= 0) atomicAdd(sm.hot + hj * SDE_HESS_HOT + hq, x);
                        else atomicAdd(hess + (size_t)my_col * p + __ldg(cols + j), x);
                    }
                })
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < hh.n * hh.n; i += SDE_NT) {
        const int r = i / hh.n, c = i % hh.n;
        const double x = sm.hot[r * SDE_HESS_HOT + c];
        if (x != 0.0) atomicAdd(hess + (size_t)hh.cols[c] * p + hh.cols[r], x);
    }
}
#undef SSDE_SLOT_GROUPS

// max slots per warp-tile and "every warp-tile is uniform" (device-built designs are checked on the device)
__global__ void design_shape_kernel(const WtDesc* __restrict__ desc, int64_t nwt, int* __restrict__ out) {
    int smax = 0, nonuni = 0, kpmax = 0, alias = 0;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nwt; q += (int64_t)gridDim.x * blockDim.x) {
        const WtDesc d = desc[q];
        const int S = slots_of(d.kmax);
        smax = max(smax, S);
        if (d.flags & WT_ALIAS_MASK) alias = 1;
#pragma unroll
        for (int p = 0; p < MAX_NP; ++p) kpmax = max(kpmax, (int)((d.kmax >> (8 * p)) & 255u));
        if (S > 0 && !(d.flags & WT_UNIFORM)) nonuni = 1;
    }
    atomicMax(out + 0, smax);
    if (nonuni) atomicOr(out + 1, 1);
    atomicMax(out + 2, kpmax);
    if (alias) atomicOr(out + 3, 1);             // aliased parameters: not understood by sde_decay_kernel
}

// ---------------------------------------------------------------------------------------------
// Decay models (nllk_sde.hpp:47-59): column c of X_re listed in col_decay is multiplied, row by
// row, by exp(-rho_k t_decay(row)), rho_k = exp(log_decay(ind_decay)), before the linear
// predictor is formed -- the design then depends on a parameter, so the values cannot be
// pre-scaled.  Same pass as sde_fused_kernel with the factor applied per nonzero; the transposed
// product and d nllk / d log_decay_k = sum eta_bar * x * theta_c * fac * (-rho_k t) are
// accumulated with atomics (decay models are small: R/sde.R:162-177 builds them for single
// experiments, not for the 1e8-row workloads).
// ---------------------------------------------------------------------------------------------
struct DecayArgs {
    const double* t_decay;     // [n_par, n_pad] permuted planes of other_data$t_decay
    const int32_t* dec_of_col; // [p_theta]: decay index of a theta column, or -1
    const double* par;         // joint parameter vector (log_decay at o_dec)
    const double* par_dot;     // tangent pass: direction
    int o_dec, n_dec;
    double* grad_decay;        // [n_dec] (R = Dual: [2 n_dec], tangents second)
};

template <class R>
__device__ __forceinline__ void atomic_add_r(double* base, int i, int stride, const R& v) {
    atomicAdd(base + i, value(v));
    if constexpr (!std::is_same<R, double>::value) atomicAdd(base + stride + i, v.d);
}

template <int MODEL, int ND, class R = double>
__global__ void __launch_bounds__(SDE_NT) sde_decay_kernel(SdeArgs a, DecayArgs dc) {
    constexpr int NP = (MODEL == MODEL_BM) ? ND + 1 : ND + 2;
    constexpr int NWARP = SDE_NT / 32;
    constexpr int MAXDEC = 16;
    __shared__ double red[8];
    __shared__ R sgrad[SDE_SGRAD];     // per-CTA accumulators of d nllk / d theta (when p_theta fits)
    __shared__ R sdec[MAXDEC];         // ... and of d nllk / d log_decay
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    GradAccT<R> gacc{(a.p_theta <= SDE_SGRAD) ? sgrad : nullptr, a.grad_theta, a.p_theta};
    for (int i = tid; i < SDE_SGRAD; i += SDE_NT) sgrad[i] = 0.0;
    if (tid < MAXDEC) sdec[tid] = 0.0;
    __syncthreads();
    R rho[MAXDEC];
#pragma unroll
    for (int k = 0; k < MAXDEC; ++k)
        rho[k] = (k < dc.n_dec) ? exp(ScalarOf<R>::make(dc.par[dc.o_dec + k], dc.par_dot ? dc.par_dot[dc.o_dec + k] : 0.0)) : R(0.0);
    auto rho_of = [&](int k) {
        R r = 0.0;
#pragma unroll
        for (int j = 0; j < MAXDEC; ++j) if (j == k) r = rho[j];
        return r;
    };
    R llk = 0.0;
    for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int64_t q = tile * NWARP + warp;
        const int64_t base = q * WT + lane;
        const int64_t row0 = q * WT + (int64_t)lane * LC;
        const WtDesc d = a.X.desc[q];
        const int S = slots_of(d.kmax);
        const bool uniform = (d.flags & WT_UNIFORM) != 0;
        const double* vals = a.X.val + d.val_off + lane;
        const uint32_t* cols = a.X.col + d.col_off + (uniform ? 0 : lane);
#pragma unroll 1
        for (int k = 0; k < LC; ++k) {
            const int64_t pos = base + k * 32;
            const uint8_t f0 = a.flags[pos];
            const bool live = f0 != 0xff && !(f0 & ROW_LAST);      // ID(i) == ID(i+1), nllk_sde.hpp:79
            const int64_t npos = (k < LC - 1) ? pos + 32 : row_pos(row0 + k + 1);
            // nonzero j of parameter p: value, column, decay factor (all lanes walk the same slots)
            auto each = [&](auto&& fn) {
                int j = 0;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const int kp = (int)((d.kmax >> (8 * p)) & 255u);
                    for (int jj = 0; jj < kp; ++jj, ++j) {
                        const double x = vals[(size_t)(k * S + j) * 32];
                        const uint32_t c = uniform ? cols[j] : cols[(size_t)(k * S + j) * 32];
                        const int dk = dc.dec_of_col[c];
                        R fac = 1.0, dfac = 0.0;                 // factor and d factor / d log_decay
                        if (dk >= 0) {
                            const double t = dc.t_decay[(size_t)p * a.X.n_pad + pos];
                            const R r = rho_of(dk);
                            fac = exp(-r * t);                   // nllk_sde.hpp:53
                            dfac = -r * t * fac;
                        }
                        fn(p, x, c, dk, fac, dfac);
                    }
                }
            };
            R eta[NP], eb[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) { eta[p] = 0.0; eb[p] = 0.0; }
            each([&](int p, double x, uint32_t c, int, const R& fac, const R&) {
#pragma unroll
                for (int pp = 0; pp < NP; ++pp) if (pp == p) eta[pp] += x * (theta_at<R>(a.theta, c) * fac);
            });
            if (live) {
                const uint8_t f1 = a.flags[npos];
                sde_row<MODEL, ND>(eta, a.dt[pos], (unsigned)((f0 | f1) >> 3),
                                   [&](int dd, int nx) { return a.obs[(size_t)dd * a.X.n_pad + (nx ? npos : pos)]; }, llk, eb);
            }
            if (a.want_grad) {
                each([&](int p, double x, uint32_t c, int dk, const R& fac, const R& dfac) {
                    R e = 0.0;
#pragma unroll
                    for (int pp = 0; pp < NP; ++pp) if (pp == p) e = eb[pp];
                    R gt = e * (x * fac);
                    R gd = (dk >= 0) ? R(e * (x * (theta_at<R>(a.theta, c) * dfac))) : R(0.0);
                    if (uniform) {
                        // every lane holds the same column: one shared-memory atomic per warp
                        gt = warp_sum(gt);
                        if (lane == 0 && nonzero(gt)) grad_add(gacc, c, gt);
                        if (dk >= 0) {                           // dk is warp-uniform here
                            gd = warp_sum(gd);
                            if (lane == 0) { atomicAdd(&reinterpret_cast<double*>(&sdec[dk])[0], value(gd));
                                             if constexpr (!std::is_same<R, double>::value) atomicAdd(&sdec[dk].d, gd.d); }
                        }
                    } else {
                        if (nonzero(gt)) grad_add(gacc, c, gt);
                        if (dk >= 0 && nonzero(gd)) { atomicAdd(&reinterpret_cast<double*>(&sdec[dk])[0], value(gd));
                                                      if constexpr (!std::is_same<R, double>::value) atomicAdd(&sdec[dk].d, gd.d); }
                    }
                });
            }
        }
    }
    __syncthreads();
    if (a.want_grad) {
        grad_flush(gacc, SDE_NT);
        if (tid < dc.n_dec) atomic_add_r<R>(dc.grad_decay, tid, dc.n_dec, sdec[tid]);
    }
    const double bl = block_sum<SDE_NT>(value(llk), red);
    if (threadIdx.x == 0) a.block_llk[blockIdx.x] = bl;
}

}  // namespace ssde
