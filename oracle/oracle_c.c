/* CPU oracle in plain C -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 *
 * A restatement of the reference objective with the SAME dense algebra the TMB template uses
 * (2d x 2d state, d x d innovation covariance, F.inverse(), T - K Z, T P L' + Q ...) plus a
 * hand-written reverse-mode sweep of that dense recursion (what TMB's tape replay computes).
 * It is deliberately NOT the decoupled / time-parallel formulation of the CUDA engine, so the
 * two are independent derivations of the same numbers.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from this file.  The product path never links or calls it.
 *
 * PARITY UNPINNED w.r.t. TMB (no R / TMB in this image, no numeric fixture in the reference);
 * pinned by tests/test_oracle.py against oracle_np.py, dense-MVN and scipy known answers.
 *
 * Reference files followed (relative to /root/reference):
 *   src/nllk/nllk_ctcrw.hpp:30-91   makeH/makeT/makeQ/makeB_ctcrw
 *   src/nllk/nllk_ctcrw.hpp:181-247 Kalman loop
 *   src/nllk/nllk_sde.hpp:77-84 + src/nllk/tr_dens.hpp:32-37,45-52   BM / OU transition sums
 *   src/nllk/nllk_ctcrw.hpp:143-156 linear predictor + natural-scale transform
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MS 6 /* max state dimension 2*d, d <= 3 */
#define MO 3

typedef struct { double v[MS][MS]; } mat;

static void mzero(mat* A) { memset(A, 0, sizeof(mat)); }
/* C = A(ra x ca) * B(ca x cb) */
static void mmul(const mat* A, const mat* B, int ra, int ca, int cb, mat* C) {
    mat R; mzero(&R);
    for (int i = 0; i < ra; ++i) for (int k = 0; k < ca; ++k) { double a = A->v[i][k]; if (a == 0.0) continue; for (int j = 0; j < cb; ++j) R.v[i][j] += a * B->v[k][j]; }
    *C = R;
}
static void mtrans(const mat* A, int r, int c, mat* T) { mat R; mzero(&R); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) R.v[j][i] = A->v[i][j]; *T = R; }
static void madd(mat* A, const mat* B, int r, int c, double s) { for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) A->v[i][j] += s * B->v[i][j]; }

/* inverse and determinant of a d x d matrix (d <= 3), det as in nllk_ctcrw.hpp:12-24 */
static double minv(const mat* F, int d, mat* Fi) {
    mzero(Fi);
    if (d == 1) { Fi->v[0][0] = 1.0 / F->v[0][0]; return F->v[0][0]; }
    if (d == 2) {
        double det = F->v[0][0] * F->v[1][1] - F->v[1][0] * F->v[0][1];
        Fi->v[0][0] = F->v[1][1] / det; Fi->v[0][1] = -F->v[0][1] / det;
        Fi->v[1][0] = -F->v[1][0] / det; Fi->v[1][1] = F->v[0][0] / det;
        return det;
    }
    const double (*a)[MS] = F->v;
    double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1], c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2], c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    double det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
    Fi->v[0][0] = c00 / det; Fi->v[1][0] = c01 / det; Fi->v[2][0] = c02 / det;
    Fi->v[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) / det;
    Fi->v[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) / det;
    Fi->v[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) / det;
    Fi->v[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) / det;
    Fi->v[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) / det;
    Fi->v[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) / det;
    return det;
}

/* makeT / makeQ / makeB, nllk_ctcrw.hpp:45-91 */
static void make_TQB(double beta, double sigma, double dt, int d, mat* T, mat* Q, mat* B) {
    mzero(T); mzero(Q); mzero(B);
    double e1 = exp(-beta * dt), e2 = exp(-2 * beta * dt);
    for (int i = 0; i < d; ++i) {
        T->v[2 * i][2 * i] = 1;
        T->v[2 * i][2 * i + 1] = (1 - e1) / beta;
        T->v[2 * i + 1][2 * i + 1] = e1;
        Q->v[2 * i][2 * i] = (sigma / beta) * (sigma / beta) * (dt - 2 / beta * (1 - e1) + 1 / (2 * beta) * (1 - e2));
        Q->v[2 * i][2 * i + 1] = sigma * sigma / (2 * beta * beta) * (1 - 2 * e1 + e2);
        Q->v[2 * i + 1][2 * i] = Q->v[2 * i][2 * i + 1];
        Q->v[2 * i + 1][2 * i + 1] = sigma * sigma / (2 * beta) * (1 - e2);
        B->v[2 * i][i] = dt - (1 - e1) / beta;
        B->v[2 * i + 1][i] = 1 - e1;
    }
}

/* ---------------------------------------------------------------------------------------------
 * CTCRW: data part of the nllk (no penalty) and, if par_bar != NULL, its adjoint w.r.t. the
 * linear predictors par_mat (n x (d+2), COLUMN-major like par_vec: par[j*n + i]) and
 * log_sigma_obs.  obs is column-major n x d with NaN = missing; a0 is n_ID x 2d column-major;
 * P0 is 2d x 2d column-major.  Tracks are processed independently (OpenMP over tracks when
 * nthreads > 1).  Returns nllk (data part).
 * ------------------------------------------------------------------------------------------- */
double oracle_ctcrw(int64_t n, int d, const double* ID, const double* times, const double* obs,
                    const double* par, const double* a0, int64_t n_ID, const double* P0,
                    double log_sigma_obs, double* par_bar, double* g_log_sigma_obs, double* aest_all,
                    int nthreads) {
    const int m = 2 * d, np_ = d + 2;
    /* track boundaries */
    int64_t* ts = (int64_t*)malloc(sizeof(int64_t) * (n + 1));
    int64_t nt = 0;
    ts[nt++] = 0;
    for (int64_t i = 1; i < n; ++i) if (ID[i] != ID[i - 1]) ts[nt++] = i;
    ts[nt] = n;
    (void)n_ID;
    const double sigma_obs = exp(log_sigma_obs);
    double total = 0.0, gsig = 0.0;
    if (par_bar) memset(par_bar, 0, sizeof(double) * n * np_);
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : total, gsig)
#endif
    for (int64_t k = 0; k < nt; ++k) {
        const int64_t s = ts[k], e = ts[k + 1], len = e - s;
        double* sa = par_bar ? (double*)malloc(sizeof(double) * len * m) : NULL;       /* aest before row */
        double* sP = par_bar ? (double*)malloc(sizeof(double) * len * m * m) : NULL;   /* Pest before row */
        double a[MS]; mat P; mzero(&P);
        for (int r = 0; r < m; ++r) { a[r] = a0[(size_t)r * n_ID + k]; for (int c = 0; c < m; ++c) P.v[r][c] = P0[(size_t)c * m + r]; }
        mat Z; mzero(&Z); for (int i = 0; i < d; ++i) Z.v[i][2 * i] = 1;
        mat Zt; mtrans(&Z, d, m, &Zt);
        mat H; mzero(&H); for (int i = 0; i < d; ++i) H.v[i][i] = sigma_obs * sigma_obs;
        double llk = 0.0;
        if (aest_all) for (int r = 0; r < m; ++r) aest_all[(size_t)r * n + s] = a[r];
        for (int64_t i = s + 1; i < e; ++i) {
            /* dtimes(i) = times(i+1) - times(i), dtimes(n-1) = 1 (:126-129).  On the last row of a
             * track the reference uses the cross-track difference; that prediction is discarded
             * at the next iteration (:196-200), so dt = 1 is used there to keep it finite. */
            const double dt = (i + 1 < e) ? times[i + 1] - times[i] : 1.0;
            const double tau = exp(par[(size_t)d * n + i]), nu = exp(par[(size_t)(d + 1) * n + i]);
            const double beta = 1 / tau, sigma = 2 * nu / sqrt(M_PI * tau);       /* :152-156 */
            mat T, Q, B; make_TQB(beta, sigma, dt, d, &T, &Q, &B);
            double Bmu[MS];
            for (int r = 0; r < m; ++r) { double t = 0; for (int c = 0; c < d; ++c) t += B.v[r][c] * par[(size_t)c * n + i]; Bmu[r] = t; }
            if (sa) { memcpy(sa + (i - s) * m, a, sizeof(double) * m); for (int r = 0; r < m; ++r) memcpy(sP + ((i - s) * m + r) * m, P.v[r], sizeof(double) * m); }
            mat Tt; mtrans(&T, m, m, &Tt);
            double an[MS];
            if (isnan(obs[i])) {                                                   /* :214-217 */
                for (int r = 0; r < m; ++r) { double t = 0; for (int c = 0; c < m; ++c) t += T.v[r][c] * a[c]; an[r] = t + Bmu[r]; }
                mat TP; mmul(&T, &P, m, m, m, &TP); mmul(&TP, &Tt, m, m, m, &P); madd(&P, &Q, m, m, 1.0);
            } else {
                double u[MO];
                for (int r = 0; r < d; ++r) u[r] = obs[(size_t)r * n + i] - a[2 * r];             /* :221 */
                mat PZt, F, Fi; mmul(&P, &Zt, m, m, d, &PZt); mmul(&Z, &PZt, d, m, d, &F); madd(&F, &H, d, d, 1.0);
                const double detF = minv(&F, d, &Fi);
                double uFu = 0; for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) uFu += u[r] * Fi.v[c][r] * u[c];
                llk -= (log(detF) + uFu) / 2;                                      /* :234 */
                mat TP, K; mmul(&T, &P, m, m, m, &TP); mmul(&TP, &Zt, m, m, d, &K); mmul(&K, &Fi, m, d, d, &K);   /* :236 */
                for (int r = 0; r < m; ++r) { double t = 0; for (int c = 0; c < m; ++c) t += T.v[r][c] * a[c]; for (int c = 0; c < d; ++c) t += K.v[r][c] * u[c]; an[r] = t + Bmu[r]; }
                mat KZ, L, Lt; mmul(&K, &Z, m, d, m, &KZ); L = T; madd(&L, &KZ, m, m, -1.0); mtrans(&L, m, m, &Lt);
                mmul(&TP, &Lt, m, m, m, &P); madd(&P, &Q, m, m, 1.0);              /* :240-241 */
            }
            memcpy(a, an, sizeof(double) * m);
            if (aest_all) for (int r = 0; r < m; ++r) aest_all[(size_t)r * n + i] = a[r];
        }
        total += -llk;
        if (par_bar) {
            /* reverse sweep over rows e-1 .. s+1 */
            double ab[MS]; memset(ab, 0, sizeof(ab));
            mat Pb; mzero(&Pb);
            double Hb_tr = 0.0;
            for (int64_t i = e - 1; i > s; --i) {
                const double dt = (i + 1 < e) ? times[i + 1] - times[i] : 1.0;
                const double tau = exp(par[(size_t)d * n + i]), nu = exp(par[(size_t)(d + 1) * n + i]);
                const double beta = 1 / tau, sigma = 2 * nu / sqrt(M_PI * tau);
                mat T, Q, B; make_TQB(beta, sigma, dt, d, &T, &Q, &B);
                mat Tt; mtrans(&T, m, m, &Tt);
                double ap[MS]; mat P; mzero(&P);
                memcpy(ap, sa + (i - s) * m, sizeof(double) * m);
                for (int r = 0; r < m; ++r) memcpy(P.v[r], sP + ((i - s) * m + r) * m, sizeof(double) * m);
                if (i == e - 1) { memset(ab, 0, sizeof(ab)); mzero(&Pb); }          /* prediction discarded */
                mat Tb, Qb, Bb; mzero(&Tb); mzero(&Bb); Qb = Pb;
                double a_in[MS]; mat P_in; mzero(&P_in); memset(a_in, 0, sizeof(a_in));
                /* a+ = T a + [K u] + B mu */
                for (int r = 0; r < m; ++r) for (int c = 0; c < m; ++c) { Tb.v[r][c] += ab[r] * ap[c]; a_in[c] += T.v[r][c] * ab[r]; }
                for (int r = 0; r < m; ++r) for (int c = 0; c < d; ++c) { Bb.v[r][c] += ab[r] * par[(size_t)c * n + i]; par_bar[(size_t)c * n + i] += B.v[r][c] * ab[r]; }
                if (isnan(obs[i])) {
                    /* P+ = T P T' + Q */
                    mat t1, t2, Pbt; mtrans(&Pb, m, m, &Pbt);
                    mmul(&Tt, &Pb, m, m, m, &t1); mmul(&t1, &T, m, m, m, &P_in);
                    mat Pt; mtrans(&P, m, m, &Pt);
                    mmul(&Pb, &T, m, m, m, &t1); mmul(&t1, &Pt, m, m, m, &t2); madd(&Tb, &t2, m, m, 1.0);
                    mmul(&Pbt, &T, m, m, m, &t1); mmul(&t1, &P, m, m, m, &t2); madd(&Tb, &t2, m, m, 1.0);
                } else {
                    double u[MO];
                    for (int r = 0; r < d; ++r) u[r] = obs[(size_t)r * n + i] - ap[2 * r];
                    mat Z; mzero(&Z); for (int q = 0; q < d; ++q) Z.v[q][2 * q] = 1;
                    mat Zt; mtrans(&Z, d, m, &Zt);
                    mat H; mzero(&H); for (int q = 0; q < d; ++q) H.v[q][q] = sigma_obs * sigma_obs;
                    mat PZt, F, Fi; mmul(&P, &Zt, m, m, d, &PZt); mmul(&Z, &PZt, d, m, d, &F); madd(&F, &H, d, d, 1.0);
                    minv(&F, d, &Fi);
                    mat TP, TPZt, K; mmul(&T, &P, m, m, m, &TP); mmul(&TP, &Zt, m, m, d, &TPZt); mmul(&TPZt, &Fi, m, d, d, &K);
                    mat KZ, L; mmul(&K, &Z, m, d, m, &KZ); L = T; madd(&L, &KZ, m, m, -1.0);
                    mat Pt; mtrans(&P, m, m, &Pt);
                    /* P+ = T P L' + Q :  Tb += Pb L P' ; P_in += T' Pb L ; Lb = Pb' T P */
                    mat t1, t2, Lb, Pbt; mtrans(&Pb, m, m, &Pbt);
                    mmul(&Pb, &L, m, m, m, &t1); mmul(&t1, &Pt, m, m, m, &t2); madd(&Tb, &t2, m, m, 1.0);
                    mmul(&Tt, &Pb, m, m, m, &t1); mmul(&t1, &L, m, m, m, &t2); madd(&P_in, &t2, m, m, 1.0);
                    mmul(&Pbt, &TP, m, m, m, &Lb);
                    /* L = T - K Z */
                    madd(&Tb, &Lb, m, m, 1.0);
                    mat Kb; mmul(&Lb, &Zt, m, m, d, &Kb); for (int r = 0; r < m; ++r) for (int c = 0; c < d; ++c) Kb.v[r][c] = -Kb.v[r][c];
                    /* a+ : K u */
                    double ub[MO]; memset(ub, 0, sizeof(ub));
                    for (int r = 0; r < m; ++r) for (int c = 0; c < d; ++c) { Kb.v[r][c] += ab[r] * u[c]; ub[c] += K.v[r][c] * ab[r]; }
                    /* K = T P Z' Fi */
                    mat Fit; mtrans(&Fi, d, d, &Fit);
                    mat KbFit; mmul(&Kb, &Fit, m, d, d, &KbFit);                 /* m x d */
                    mat PZtT; mtrans(&PZt, m, d, &PZtT);                           /* d x m = Z P' */
                    mmul(&KbFit, &PZtT, m, d, m, &t1); madd(&Tb, &t1, m, m, 1.0);  /* Tb += Kb Fi' (P Z')' */
                    mmul(&Tt, &KbFit, m, m, d, &t1); mmul(&t1, &Z, m, d, m, &t2); madd(&P_in, &t2, m, m, 1.0);   /* P_in += T' Kb Fi' Z */
                    mat TPZtT, Fib; mtrans(&TPZt, m, d, &TPZtT); mmul(&TPZtT, &Kb, d, m, d, &Fib);    /* Fib = (T P Z')' Kb */
                    /* nll = (log det F + u' Fi' u)/2 */
                    for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) { ub[r] += 0.5 * (Fi.v[r][c] + Fi.v[c][r]) * u[c]; Fib.v[r][c] += 0.5 * u[r] * u[c]; }
                    mat Fb; mzero(&Fb); madd(&Fb, &Fit, d, d, 0.5);
                    /* Fi = F^-1 : Fb -= Fi' Fib Fi' */
                    mmul(&Fit, &Fib, d, d, d, &t1); mmul(&t1, &Fit, d, d, d, &t2); madd(&Fb, &t2, d, d, -1.0);
                    /* F = Z P Z' + H */
                    mmul(&Zt, &Fb, m, d, d, &t1); mmul(&t1, &Z, m, d, m, &t2); madd(&P_in, &t2, m, m, 1.0);
                    for (int q = 0; q < d; ++q) Hb_tr += Fb.v[q][q];
                    /* u = y - Z a */
                    for (int q = 0; q < d; ++q) a_in[2 * q] -= ub[q];
                }
                /* T, Q, B -> beta, sigma (derivatives of makeT/makeQ/makeB) */
                {
                    const double e1 = exp(-beta * dt), e2 = exp(-2 * beta * dt);
                    double bb = 0.0, sb = 0.0;
                    for (int q = 0; q < d; ++q) {
                        const double T12b = Tb.v[2 * q][2 * q + 1], T22b = Tb.v[2 * q + 1][2 * q + 1];
                        const double Q11b = Qb.v[2 * q][2 * q], Q12b = Qb.v[2 * q][2 * q + 1] + Qb.v[2 * q + 1][2 * q], Q22b = Qb.v[2 * q + 1][2 * q + 1];
                        const double B1b = Bb.v[2 * q][q], B2b = Bb.v[2 * q + 1][q];
                        /* T12 = (1-e1)/beta */
                        const double dT12 = dt * e1 / beta - (1 - e1) / (beta * beta);
                        const double dT22 = -dt * e1;
                        /* Q11 = sigma^2/beta^2 * g, g = dt - 2(1-e1)/beta + (1-e2)/(2 beta) */
                        const double g = dt - 2 / beta * (1 - e1) + 1 / (2 * beta) * (1 - e2);
                        const double dg = 2 * (1 - e1) / (beta * beta) - 2 * dt * e1 / beta - (1 - e2) / (2 * beta * beta) + dt * e2 / beta;
                        const double s2 = sigma * sigma;
                        const double dQ11 = -2 * s2 / (beta * beta * beta) * g + s2 / (beta * beta) * dg;
                        const double q12 = 1 - 2 * e1 + e2;
                        const double dQ12 = -s2 / (beta * beta * beta) * q12 + s2 / (2 * beta * beta) * (2 * dt * e1 - 2 * dt * e2);
                        const double dQ22 = -s2 / (2 * beta * beta) * (1 - e2) + s2 / (2 * beta) * (2 * dt * e2);
                        const double dB1 = -dT12, dB2 = dt * e1;
                        bb += T12b * dT12 + T22b * dT22 + Q11b * dQ11 + Q12b * dQ12 + Q22b * dQ22 + B1b * dB1 + B2b * dB2;
                        sb += (Q11b * Q.v[2 * q][2 * q] + Q12b * Q.v[2 * q][2 * q + 1] + Q22b * Q.v[2 * q + 1][2 * q + 1]) * 2 / sigma;
                    }
                    /* beta = 1/tau, sigma = 2 nu / sqrt(pi tau), tau = exp(eta_tau), nu = exp(eta_nu) */
                    par_bar[(size_t)d * n + i] += bb * (-beta) + sb * (-0.5 * sigma);
                    par_bar[(size_t)(d + 1) * n + i] += sb * sigma;
                }
                memcpy(ab, a_in, sizeof(ab));
                Pb = P_in;
            }
            gsig += 2 * sigma_obs * sigma_obs * Hb_tr;
            free(sa); free(sP);
        }
    }
    free(ts);
    if (g_log_sigma_obs) *g_log_sigma_obs = gsig;
    return total;
}

/* ---------------------------------------------------------------------------------------------
 * BM / OU (nllk_sde.hpp:77-84 + tr_dens.hpp): data part of the nllk and adjoint w.r.t. par.
 * model 0 = BM, 1 = OU.  par column-major n x n_par.
 * ------------------------------------------------------------------------------------------- */
double oracle_sde(int model, int64_t n, int d, const double* ID, const double* times, const double* obs,
                  const double* par, double* par_bar, int nthreads) {
    const int np_ = (model == 0) ? d + 1 : d + 2;
    if (par_bar) memset(par_bar, 0, sizeof(double) * n * np_);
    double llk = 0.0;
    const double LS2P = 0.5 * log(2 * M_PI);
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads) reduction(+ : llk)
#endif
    for (int64_t i = 1; i < n; ++i) {
        if (ID[i - 1] != ID[i]) continue;
        const double dt = times[i] - times[i - 1];
        const int64_t j = i - 1;                               /* parameters of row i-1 */
        for (int q = 0; q < d; ++q) {
            const double z0 = obs[(size_t)q * n + j], z1 = obs[(size_t)q * n + i];
            if (isnan(z0) || isnan(z1)) continue;             /* tr_dens.hpp:31 */
            if (model == 0) {
                const double mean = z0 + par[(size_t)q * n + j] * dt;
                const double sd = exp(par[(size_t)d * n + j]) * sqrt(dt);
                const double r = (z1 - mean) / sd;
                llk += -LS2P - log(sd) - 0.5 * r * r;
                if (par_bar) { par_bar[(size_t)q * n + j] += -r * dt / sd; par_bar[(size_t)d * n + j] += 1 - r * r; }
            } else {
                const double mu = par[(size_t)q * n + j], tau = exp(par[(size_t)d * n + j]), kappa = exp(par[(size_t)(d + 1) * n + j]);
                const double ph = exp(-dt / tau);
                const double mean = mu + ph * (z0 - mu);
                const double var = kappa * (1 - exp(-2 * dt / tau));
                const double sd = sqrt(var);
                const double r = (z1 - mean) / sd;
                llk += -LS2P - log(sd) - 0.5 * r * r;
                if (par_bar) {
                    const double dph = ph * dt / tau;
                    const double dvar = -kappa * 2 * exp(-2 * dt / tau) * dt / tau;
                    par_bar[(size_t)q * n + j] += -r * (1 - ph) / sd;
                    par_bar[(size_t)d * n + j] += 0.5 * dvar / var * (1 - r * r) - r / sd * dph * (z0 - mu);
                    par_bar[(size_t)(d + 1) * n + j] += 0.5 * (1 - r * r);
                }
            }
        }
    }
    return -llk;
}

/* CSR products for the linear predictor (stacked [X_fe | X_re], (n_par n) x p, standard CSR) */
void oracle_spmv(int64_t nrow, const int64_t* rowptr, const int32_t* col, const double* val,
                 const double* theta, double* out, int nthreads) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads < 1 ? 1 : nthreads)
#endif
    for (int64_t r = 0; r < nrow; ++r) {
        double s = 0.0;
        for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) s += val[k] * theta[col[k]];
        out[r] = s;
    }
}
void oracle_spmv_t(int64_t nrow, int64_t ncol, const int64_t* rowptr, const int32_t* col,
                   const double* val, const double* x, double* out) {
    memset(out, 0, sizeof(double) * ncol);
    for (int64_t r = 0; r < nrow; ++r) {
        const double xr = x[r];
        if (xr == 0.0) continue;
        for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) out[col[k]] += val[k] * xr;
    }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
