// Exact-transition CTCRW simulator on the device, one thread per (track, dimension), sequential
// in time.  Follows SDE$simulate for type "CTCRW" (R/sde.R:1448-1478) and CTCRW_cov
// (R/utility.R:188-196): the transition from row i-1 to row i uses the parameters of row i-1.
// Only used to make synthetic tracks of the benchmark shapes; not on the likelihood path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ssde {

__global__ void ctcrw_sim_kernel(int64_t n_tracks, int64_t m, const double* __restrict__ times,
                                 const double* __restrict__ tau, const double* __restrict__ nu,
                                 const double* __restrict__ mu, const double* __restrict__ e1,
                                 const double* __restrict__ e2, double* __restrict__ z) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    const int64_t base = t * m;
    double v = 0.0, zc = z[base];
    for (int64_t i = 1; i < m; ++i) {
        const int64_t j = base + i - 1;
        const double ta = tau[j], b = 1.0 / ta, s = 2.0 * nu[j] / sqrt(3.14159265358979323846 * ta);
        const double d = times[j + 1] - times[j];
        const double mm = mu ? mu[j] : 0.0;
        const double p = exp(-b * d), p2 = exp(-2.0 * b * d);
        const double mean_v = p * v + (1.0 - p) * mm;
        const double mean_z = zc + mm * d + (v - mm) / b * (1.0 - p);
        const double qvv = s * s / (2.0 * b) * (1.0 - p2);
        const double qzz = (s / b) * (s / b) * (d + (1.0 - p2) / (2.0 * b) - 2.0 * (1.0 - p) / b);
        const double qvz = s * s / (2.0 * b * b) * (1.0 - 2.0 * p + p2);
        const double l11 = sqrt(qvv), l21 = qvz / l11;
        const double l22 = sqrt(fmax(qzz - l21 * l21, 0.0));
        v = mean_v + l11 * e1[j + 1];
        zc = mean_z + l21 * e1[j + 1] + l22 * e2[j + 1];
        z[j + 1] = zc;
    }
}

// Exact-transition OU simulator (SDE$simulate for type "OU", R/sde.R:1439-1447): one thread per
// track, z_i ~ N(mu + exp(-dt/tau)(z_{i-1} - mu), kappa (1 - exp(-2 dt/tau))) with the parameters
// of row i-1.
__global__ void ou_sim_kernel(int64_t n_tracks, int64_t m, const double* __restrict__ times,
                              const double* __restrict__ mu, const double* __restrict__ tau,
                              const double* __restrict__ kappa, const double* __restrict__ e,
                              double* __restrict__ z) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    const int64_t base = t * m;
    double zc = z[base];
    for (int64_t i = 1; i < m; ++i) {
        const int64_t j = base + i - 1;
        const double d = times[j + 1] - times[j];
        const double p = exp(-d / tau[j]);
        const double mean = p * zc + (1.0 - p) * mu[j];
        const double sd = sqrt(kappa[j] * (1.0 - p * p));
        zc = mean + sd * e[j + 1];
        z[j + 1] = zc;
    }
}

}  // namespace ssde
