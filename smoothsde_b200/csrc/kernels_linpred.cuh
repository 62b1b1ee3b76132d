// Linear predictors from the sparse spline design matrices and their transpose products, plus
// the fused BM / OU transition-density kernel.
//
// Replaces:  par_vec = X_fe*coeff_fe + X_re*coeff_re and its reshape (nllk_ctcrw.hpp:143-149,
// nllk_sde.hpp:61-67), the natural-scale transform (nllk_ctcrw.hpp:152-156), the BM / OU
// transition densities (tr_dens.hpp:32-37, :45-52) summed over rows (nllk_sde.hpp:77-84) and
// TMB's reverse sweep through all of these (X_fe' eta_bar, X_re' eta_bar).
//
// Device layout of the design ("observation-major packed CSR", built once in ssde_create):
// X_fe and X_re are block-diagonal by parameter (rows j*n..(j+1)*n-1 belong to parameter j,
// R/sde.R:443-447), so for observation row i the nonzeros of all its n_par parameter rows are
// stored contiguously, parameter-sorted, with one packed count word per row:
//   rowptr[n+1] (u32)   start of row i's nonzeros
//   cnt[n]      (u32)   byte p = number of nonzeros of parameter p in row i  (n_par <= 4)
//   col[nnz]    (u32)   column in theta = [coeff_fe | coeff_re]
//   val[nnz]    (f64)
// One CTA stages the nonzeros of a tile of rows in shared memory with coalesced loads; one
// thread then owns one row.
#pragma once

#include "common.cuh"

namespace ssde {

constexpr int LP_NT = 128;        // threads per CTA = rows per tile
constexpr int LP_CAP = 3584;      // staged nonzeros per tile (28 per row on average)
constexpr int LP_CACHE = 64;      // per-warp gradient cache slots
constexpr double LOG_SQRT_2PI = 0.918938533204672741780329736406;

struct Design {
    int64_t n;
    int n_par;
    const uint32_t* rowptr;
    const uint32_t* cnt;
    const uint32_t* col;
    const double* val;
};

struct LpSmem {
    double val[LP_CAP];
    uint32_t col[LP_CAP];
    double cval[LP_NT / 32][LP_CACHE];
    uint32_t ctag[LP_NT / 32][LP_CACHE];
    double red[8];
};

// Stage the nonzeros of rows [r0, r1) into shared memory.  Returns false (nothing staged) if the
// tile holds more than LP_CAP nonzeros; callers then read global memory directly.
__device__ __forceinline__ bool stage_nnz(LpSmem& sm, const Design& X, int64_t r0, int64_t r1,
                                          uint32_t& s) {
    s = X.rowptr[r0];
    const uint32_t e = X.rowptr[r1];
    const uint32_t nn = e - s;
    if (nn > (uint32_t)LP_CAP) return false;
    for (uint32_t k = threadIdx.x; k < nn; k += LP_NT) {
        sm.val[k] = __ldg(X.val + s + k);
        sm.col[k] = __ldg(X.col + s + k);
    }
    return true;
}

// eta[p] = sum over the row's nonzeros of parameter p
template <int NP, bool STAGED>
__device__ __forceinline__ void row_dot(const LpSmem& sm, const Design& X, uint32_t rp, uint32_t s,
                                        uint32_t c, const double* __restrict__ theta, double* eta) {
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const uint32_t k = (c >> (8 * p)) & 255u;
        double acc = 0.0;
        for (uint32_t j = 0; j < k; ++j) {
            const double v = STAGED ? sm.val[rp - s + j] : __ldg(X.val + rp + j);
            const uint32_t cc = STAGED ? sm.col[rp - s + j] : __ldg(X.col + rp + j);
            acc = fma(v, __ldg(theta + cc), acc);
        }
        eta[p] = acc;
        rp += k;
    }
}

// ---------------------------------------------------------------------------------------------
// per-warp gradient cache: lane 0 owns LP_CACHE direct-mapped (column -> partial sum) slots in
// shared memory; a conflicting column evicts the resident one with a global atomicAdd.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cache_init(LpSmem& sm) {
    for (int i = threadIdx.x; i < (LP_NT / 32) * LP_CACHE; i += LP_NT) {
        (&sm.ctag[0][0])[i] = 0xffffffffu;
        (&sm.cval[0][0])[i] = 0.0;
    }
}
__device__ __forceinline__ void cache_add(LpSmem& sm, int warp, uint32_t c, double v, double* grad) {
    const int slot = c & (LP_CACHE - 1);
    const uint32_t tag = sm.ctag[warp][slot];
    if (tag == c) {
        sm.cval[warp][slot] += v;
    } else {
        if (tag != 0xffffffffu) atomicAdd(grad + tag, sm.cval[warp][slot]);
        sm.ctag[warp][slot] = c;
        sm.cval[warp][slot] = v;
    }
}
__device__ __forceinline__ void cache_flush(LpSmem& sm, double* grad) {
    __syncthreads();
    for (int i = threadIdx.x; i < (LP_NT / 32) * LP_CACHE; i += LP_NT) {
        const uint32_t tag = (&sm.ctag[0][0])[i];
        if (tag != 0xffffffffu) atomicAdd(grad + tag, (&sm.cval[0][0])[i]);
    }
}

// grad[col] += val * eb[p] for every nonzero of the row (whole warp participates; rows past the
// end pass c = 0).  Lanes holding the same column are summed with shuffles first.
template <int NP, bool STAGED>
__device__ __forceinline__ void row_scatter(LpSmem& sm, const Design& X, uint32_t rp, uint32_t s,
                                            uint32_t c, const double* eb, double* grad) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const uint32_t k = (c >> (8 * p)) & 255u;
        const uint32_t kmax = __reduce_max_sync(FULL, k);
        for (uint32_t j = 0; j < kmax; ++j) {
            const bool act = j < k;
            uint32_t cc = 0xffffffffu;
            double g = 0.0;
            if (act) {
                const double v = STAGED ? sm.val[rp - s + j] : __ldg(X.val + rp + j);
                cc = STAGED ? sm.col[rp - s + j] : __ldg(X.col + rp + j);
                g = v * eb[p];
            }
            const uint32_t c0 = __shfl_sync(FULL, cc, 0);
            if (__all_sync(FULL, cc == c0)) {
                if (c0 != 0xffffffffu) {
                    g = warp_sum(g);
                    if (lane == 0) cache_add(sm, warp, c0, g, grad);
                }
            } else if (act) {
                atomicAdd(grad + cc, g);
            }
        }
        rp += k;
    }
}

// ---------------------------------------------------------------------------------------------
// K1 (CTCRW):  W[i] = (mu_1..mu_ND, tau, e, s2) from eta_i = X_i theta
// ---------------------------------------------------------------------------------------------
template <int ND>
__global__ void __launch_bounds__(LP_NT) ctcrw_linpred_kernel(Design X, const double* __restrict__ theta,
                                                              const double* __restrict__ dt,
                                                              double* __restrict__ W) {
    constexpr int NP = ND + 2, NW = ND + 3;
    __shared__ LpSmem sm;
    const int64_t ntiles = (X.n + LP_NT - 1) / LP_NT;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * LP_NT;
        const int64_t r1 = (r0 + LP_NT < X.n) ? r0 + LP_NT : X.n;
        uint32_t s;
        __syncthreads();
        const bool staged = stage_nnz(sm, X, r0, r1, s);
        __syncthreads();
        const int64_t r = r0 + threadIdx.x;
        if (r < r1) {
            double eta[NP];
            const uint32_t rp = X.rowptr[r], c = X.cnt[r];
            if (staged) row_dot<NP, true>(sm, X, rp, s, c, theta, eta);
            else row_dot<NP, false>(sm, X, rp, s, c, theta, eta);
            double tau, e, s2;
            transform_row(eta[ND], eta[ND + 1], dt[r], tau, e, s2);
            double* w = W + (size_t)r * NW;
#pragma unroll
            for (int d = 0; d < ND; ++d) w[d] = eta[d];
            w[ND] = tau; w[ND + 1] = e; w[ND + 2] = s2;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1^T:  grad_theta += X' eta_bar   (eta_bar is [n, NP] row-major)
// ---------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(LP_NT) linpred_T_kernel(Design X, const double* __restrict__ eta_bar,
                                                          double* __restrict__ grad) {
    __shared__ LpSmem sm;
    cache_init(sm);
    const int64_t ntiles = (X.n + LP_NT - 1) / LP_NT;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * LP_NT;
        const int64_t r1 = (r0 + LP_NT < X.n) ? r0 + LP_NT : X.n;
        uint32_t s;
        __syncthreads();
        const bool staged = stage_nnz(sm, X, r0, r1, s);
        __syncthreads();
        const int64_t r = r0 + threadIdx.x;
        double eb[NP];
        uint32_t rp = 0, c = 0;
        if (r < r1) {
            rp = X.rowptr[r]; c = X.cnt[r];
#pragma unroll
            for (int p = 0; p < NP; ++p) eb[p] = eta_bar[(size_t)r * NP + p];
        } else {
#pragma unroll
            for (int p = 0; p < NP; ++p) eb[p] = 0.0;
        }
        if (staged) row_scatter<NP, true>(sm, X, rp, s, c, eb, grad);
        else row_scatter<NP, false>(sm, X, rp, s, c, eb, grad);
    }
    cache_flush(sm, grad);
}

// ---------------------------------------------------------------------------------------------
// fused BM / OU kernel: eta -> transition log-density of row i -> row i+1 (parameters of row i,
// dt_i = t_{i+1} - t_i; nllk_sde.hpp:80-81) -> closed-form eta_bar -> X' eta_bar, one pass.
// ---------------------------------------------------------------------------------------------
enum { MODEL_BM = 0, MODEL_OU = 1, MODEL_CTCRW = 2 };

template <int MODEL, int ND>
__global__ void __launch_bounds__(LP_NT) sde_fused_kernel(Design X, const double* __restrict__ theta,
                                                          const double* __restrict__ obs,
                                                          const double* __restrict__ dt,
                                                          const uint8_t* __restrict__ flags,
                                                          int want_grad, double* __restrict__ grad,
                                                          double* __restrict__ block_llk) {
    constexpr int NP = (MODEL == MODEL_BM) ? ND + 1 : ND + 2;
    __shared__ LpSmem sm;
    cache_init(sm);
    double llk = 0.0;
    const int64_t ntiles = (X.n + LP_NT - 1) / LP_NT;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * LP_NT;
        const int64_t r1 = (r0 + LP_NT < X.n) ? r0 + LP_NT : X.n;
        uint32_t s;
        __syncthreads();
        const bool staged = stage_nnz(sm, X, r0, r1, s);
        __syncthreads();
        const int64_t r = r0 + threadIdx.x;
        double eb[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) eb[p] = 0.0;
        uint32_t rp = 0, c = 0;
        if (r < r1) {
            rp = X.rowptr[r]; c = X.cnt[r];
            const uint8_t f0 = flags[r];
            if (!(f0 & ROW_LAST)) {                 // ID(i) == ID(i+1), nllk_sde.hpp:79
                const uint8_t f1 = flags[r + 1];
                double eta[NP];
                if (staged) row_dot<NP, true>(sm, X, rp, s, c, theta, eta);
                else row_dot<NP, false>(sm, X, rp, s, c, theta, eta);
                const double d_t = dt[r];
                if (MODEL == MODEL_BM) {
                    // mean = z0 + mu dt, sd = exp(eta_s) sqrt(dt)   (tr_dens.hpp:35-36)
                    const double sd = exp(eta[ND]) * sqrt(d_t);
                    const double isd = 1.0 / sd, lsd = log(sd);
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        if (((f0 | f1) >> (3 + d)) & 1) continue;      // tr_dens.hpp:31
                        const double z0 = obs[(size_t)r * ND + d], z1 = obs[(size_t)(r + 1) * ND + d];
                        const double res = (z1 - (z0 + eta[d] * d_t)) * isd;
                        llk += -LOG_SQRT_2PI - lsd - 0.5 * res * res;
                        eb[d] = -res * d_t * isd;
                        eb[ND] += 1.0 - res * res;
                    }
                } else {
                    // mean = mu + exp(-dt/tau)(z0 - mu), sd = sqrt(kappa (1 - exp(-2 dt/tau)))
                    const double tau = exp(eta[ND]), kappa = exp(eta[ND + 1]);
                    const double ph = exp(-d_t / tau);
                    const double var = kappa * (1.0 - ph * ph);
                    const double sd = sqrt(var), isd = 1.0 / sd, lsd = log(sd);
                    const double dph = ph * d_t / tau;                  // d ph / d eta_tau
                    const double dlv = -2.0 * kappa * ph * dph / var;   // d log var / d eta_tau
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        if (((f0 | f1) >> (3 + d)) & 1) continue;
                        const double z0 = obs[(size_t)r * ND + d], z1 = obs[(size_t)(r + 1) * ND + d];
                        const double res = (z1 - (eta[d] + ph * (z0 - eta[d]))) * isd;
                        llk += -LOG_SQRT_2PI - lsd - 0.5 * res * res;
                        eb[d] = -res * (1.0 - ph) * isd;
                        eb[ND] += 0.5 * dlv * (1.0 - res * res) - res * isd * dph * (z0 - eta[d]);
                        eb[ND + 1] += 0.5 * (1.0 - res * res);
                    }
                }
            }
        }
        if (want_grad) {
            if (staged) row_scatter<NP, true>(sm, X, rp, s, c, eb, grad);
            else row_scatter<NP, false>(sm, X, rp, s, c, eb, grad);
        }
    }
    if (want_grad) cache_flush(sm, grad);
    const double bl = block_sum<LP_NT>(llk, sm.red);
    if (threadIdx.x == 0) block_llk[blockIdx.x] = bl;
}

}  // namespace ssde
