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
        const R var = kappa * (1.0 - exp(-2.0 * d_t / tau));     // tr_dens.hpp:50-51
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
