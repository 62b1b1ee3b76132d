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
                if (MODEL == MODEL_BM) {
                    // mean = z0 + mu dt, sd = exp(eta_s) sqrt(dt)   (tr_dens.hpp:35-36)
                    const R sd = exp(eta[ND]) * sqrt(d_t);
                    const R isd = 1.0 / sd, lsd = log(sd);
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        if (((f0 | f1) >> (3 + d)) & 1) continue;      // tr_dens.hpp:31
                        const double z0 = a.obs[(size_t)d * a.X.n_pad + pos], z1 = a.obs[(size_t)d * a.X.n_pad + npos];
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
                        if (((f0 | f1) >> (3 + d)) & 1) continue;
                        const double z0 = a.obs[(size_t)d * a.X.n_pad + pos], z1 = a.obs[(size_t)d * a.X.n_pad + npos];
                        const R res = (z1 - (eta[d] + ph * (z0 - eta[d]))) * isd;
                        llk += -LOG_SQRT_2PI - lsd - 0.5 * res * res;
                        eb[d] = -res * (1.0 - ph) * isd;
                        eb[ND] += 0.5 * dlv * (1.0 - res * res) - res * isd * dph * (z0 - eta[d]);
                        eb[ND + 1] += 0.5 * (1.0 - res * res);
                    }
                }
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

}  // namespace ssde
