// Coupled ("dense") Kalman filter algebra: the full N x N covariance recursion of
// src/nllk/nllk_ctcrw.hpp:195-247 (N = 2 n_dim), nllk_ou_ssm.hpp:163-213 and nllk_bm_ssm.hpp:127-175
// (N = n_dim), needed whenever the dimensions do not decouple:
//   * a user-supplied measurement covariance per row, H_array[, , i] (nllk_ctcrw.hpp:203-205,
//     R/sde.R:593-598: Argos error ellipses), which is a general SPD n_dim x n_dim matrix, or
//   * a user-supplied P0 (R/sde.R:551-556,582-588) that is not of the default shape.
// Same time-parallel formulation as ctcrw_math.cuh (read that header first), with matrix-valued
// elements: forward (A, b, C, eta, J) with N^2 + 2N + N(N+1) scalars (44 for the 4-state CTCRW),
// adjoint (L, z, D) with N^2 + N + N(N+1)/2 (30).  T, Q and B of all three models are
// block-diagonal with one identical SPD x SPD block per dimension (SPD = states per dimension),
// which is how a step is represented here (StepBlk); Z picks the first state of each dimension.
//
// Derivation of the adjoint of one row (state (a, P) -> (a+, P+), nllk term f):
//   u = y - Z a,  F = Z P Z' + H,  w = F^-1 u,  G = P Z' F^-1,  f = (log|F| + u'w)/2,
//   af = a + G u,  Pf = (I - G Z) P (I - G Z)' + G H G' = P - G F G',  a+ = T af + c,  P+ = T Pf T' + Q
//   abar = L' abar+ - z,   Pbar = L' Pbar+ L + sym(L' abar+ z') + D,
//   L = T (I - G Z),  z = Z' w,  D = Z' Fl Z,  Fl = (F^-1 - w w')/2
//   Tbar = abar+ af' + 2 Pbar+ T Pf,  Qbar = Pbar+,  cbar = abar+,
//   Hbar = Fl - sym(G' T' abar+ w') + G' T' Pbar+ T G
// (Pbar is the symmetric full-matrix adjoint: <Pbar, dP> = sum_ij Pbar_ij dP_ij).
#pragma once

#include <math.h>

#include "dual.cuh"

namespace ssde {

template <int N>
SSDE_HD constexpr int sym_idx(int i, int j) {       // packed upper triangle, row-major
    return (i <= j) ? (i * N - i * (i - 1) / 2 + (j - i)) : (j * N - j * (j - 1) / 2 + (i - j));
}

template <int N, class R>
SSDE_HD void sym_unpack(const R* p, R (&M)[N][N]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) M[i][j] = p[sym_idx<N>(i, j)];
}
// p = (M + M')/2
template <int N, class R>
SSDE_HD void sym_pack(const R (&M)[N][N], R* p) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        p[sym_idx<N>(i, i)] = M[i][i];
#pragma unroll
        for (int j = i + 1; j < N; ++j) p[sym_idx<N>(i, j)] = 0.5 * (M[i][j] + M[j][i]);
    }
}

template <int N, class R>
SSDE_HD void mat_mul(const R (&A)[N][N], const R (&B)[N][N], R (&C)[N][N]) {       // C = A B
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            R s = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) s += A[i][k] * B[k][j];
            C[i][j] = s;
        }
}
template <int N, class R>
SSDE_HD void mat_mul_nt(const R (&A)[N][N], const R (&B)[N][N], R (&C)[N][N]) {    // C = A B'
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            R s = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) s += A[i][k] * B[j][k];
            C[i][j] = s;
        }
}
template <int N, class R>
SSDE_HD void mat_mul_tn(const R (&A)[N][N], const R (&B)[N][N], R (&C)[N][N]) {    // C = A' B
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            R s = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) s += A[k][i] * B[k][j];
            C[i][j] = s;
        }
}
template <int N, class R>
SSDE_HD void mat_vec(const R (&A)[N][N], const R* x, R* y) {                         // y = A x
#pragma unroll
    for (int i = 0; i < N; ++i) {
        R s = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) s += A[i][k] * x[k];
        y[i] = s;
    }
}
template <int N, class R>
SSDE_HD void mat_tvec(const R (&A)[N][N], const R* x, R* y) {                        // y = A' x
#pragma unroll
    for (int i = 0; i < N; ++i) {
        R s = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) s += A[k][i] * x[k];
        y[i] = s;
    }
}

// Inverse of a general N x N matrix (N <= 4) by Gauss-Jordan elimination with partial pivoting.
// Every index is a compile-time constant after unrolling, so the row exchanges are selects and
// everything stays in registers.  Used for (I + C J)^-1 of the scan combine: that matrix has
// positive eigenvalues (C, J are PSD) but is not symmetric, and its leading minors can vanish.
template <int N, class R>
SSDE_HD void inv_general(const R (&X)[N][N], R (&Xi)[N][N]) {
    R a[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) { a[i][j] = X[i][j]; Xi[i][j] = (i == j) ? 1.0 : 0.0; }
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int r = k + 1; r < N; ++r) {
            if (fabs(value(a[r][k])) > fabs(value(a[k][k]))) {
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    const R t1 = a[k][c]; a[k][c] = a[r][c]; a[r][c] = t1;
                    const R t2 = Xi[k][c]; Xi[k][c] = Xi[r][c]; Xi[r][c] = t2;
                }
            }
        }
        const R ip = 1.0 / a[k][k];
#pragma unroll
        for (int c = 0; c < N; ++c) { a[k][c] *= ip; Xi[k][c] *= ip; }
#pragma unroll
        for (int r = 0; r < N; ++r) {
            if (r == k) continue;
            const R f = a[r][k];
#pragma unroll
            for (int c = 0; c < N; ++c) { a[r][c] -= f * a[k][c]; Xi[r][c] -= f * Xi[k][c]; }
        }
    }
}

// Inverse and determinant of a symmetric positive definite D x D matrix, D <= 3, in closed form
// (the reference uses a closed-form determinant for n_dim <= 2, nllk_ctcrw.hpp:12-24, and
// Eigen's inverse(), which is closed-form for these sizes as well).
template <class R>
SSDE_HD void sym_inv_det(const R (&F)[1][1], R (&Fi)[1][1], R& det) {
    det = F[0][0];
    Fi[0][0] = 1.0 / F[0][0];
}
template <class R>
SSDE_HD void sym_inv_det(const R (&F)[2][2], R (&Fi)[2][2], R& det) {
    det = F[0][0] * F[1][1] - F[0][1] * F[1][0];
    const R id = 1.0 / det;
    Fi[0][0] = F[1][1] * id; Fi[1][1] = F[0][0] * id;
    Fi[0][1] = -F[0][1] * id; Fi[1][0] = Fi[0][1];
}
template <class R>
SSDE_HD void sym_inv_det(const R (&F)[3][3], R (&Fi)[3][3], R& det) {
    const R c00 = F[1][1] * F[2][2] - F[1][2] * F[2][1];
    const R c01 = F[1][2] * F[2][0] - F[1][0] * F[2][2];
    const R c02 = F[1][0] * F[2][1] - F[1][1] * F[2][0];
    det = F[0][0] * c00 + F[0][1] * c01 + F[0][2] * c02;
    const R id = 1.0 / det;
    Fi[0][0] = c00 * id; Fi[0][1] = c01 * id; Fi[0][2] = c02 * id;
    Fi[1][1] = (F[0][0] * F[2][2] - F[0][2] * F[2][0]) * id;
    Fi[1][2] = (F[0][2] * F[1][0] - F[0][0] * F[1][2]) * id;
    Fi[2][2] = (F[0][0] * F[1][1] - F[0][1] * F[1][0]) * id;
    Fi[1][0] = Fi[0][1]; Fi[2][0] = Fi[0][2]; Fi[2][1] = Fi[1][2];
}

// ---------------------------------------------------------------------------------------------
// step of one row in block form:  T = I_D (x) t,  Q = I_D (x) q,  c_d = b mu_d
// ---------------------------------------------------------------------------------------------
template <int SPD, class R>
struct StepBlk {
    R t[SPD][SPD], q[SPD][SPD], b[SPD];
};

// Y = (I (x) t) X
template <int D, int SPD, class R>
SSDE_HD void blk_left(const R (&t)[SPD][SPD], const R (&X)[D * SPD][D * SPD], R (&Y)[D * SPD][D * SPD]) {
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
        for (int i = 0; i < SPD; ++i)
#pragma unroll
            for (int c = 0; c < D * SPD; ++c) {
                R s = 0.0;
#pragma unroll
                for (int j = 0; j < SPD; ++j) s += t[i][j] * X[d * SPD + j][c];
                Y[d * SPD + i][c] = s;
            }
}
// Y = (I (x) t)' X
template <int D, int SPD, class R>
SSDE_HD void blk_left_t(const R (&t)[SPD][SPD], const R (&X)[D * SPD][D * SPD], R (&Y)[D * SPD][D * SPD]) {
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
        for (int j = 0; j < SPD; ++j)
#pragma unroll
            for (int c = 0; c < D * SPD; ++c) {
                R s = 0.0;
#pragma unroll
                for (int i = 0; i < SPD; ++i) s += t[i][j] * X[d * SPD + i][c];
                Y[d * SPD + j][c] = s;
            }
}
// Y = X (I (x) t)'
template <int D, int SPD, class R>
SSDE_HD void blk_right_t(const R (&X)[D * SPD][D * SPD], const R (&t)[SPD][SPD], R (&Y)[D * SPD][D * SPD]) {
#pragma unroll
    for (int r = 0; r < D * SPD; ++r)
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
            for (int i = 0; i < SPD; ++i) {
                R s = 0.0;
#pragma unroll
                for (int j = 0; j < SPD; ++j) s += X[r][d * SPD + j] * t[i][j];
                Y[r][d * SPD + i] = s;
            }
}
// Y = X (I (x) t)
template <int D, int SPD, class R>
SSDE_HD void blk_right(const R (&X)[D * SPD][D * SPD], const R (&t)[SPD][SPD], R (&Y)[D * SPD][D * SPD]) {
#pragma unroll
    for (int r = 0; r < D * SPD; ++r)
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
            for (int j = 0; j < SPD; ++j) {
                R s = 0.0;
#pragma unroll
                for (int i = 0; i < SPD; ++i) s += X[r][d * SPD + i] * t[i][j];
                Y[r][d * SPD + j] = s;
            }
}
// y = (I (x) t) x ;  y = (I (x) t)' x
template <int D, int SPD, class R>
SSDE_HD void blk_vec(const R (&t)[SPD][SPD], const R* x, R* y) {
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
        for (int i = 0; i < SPD; ++i) {
            R s = 0.0;
#pragma unroll
            for (int j = 0; j < SPD; ++j) s += t[i][j] * x[d * SPD + j];
            y[d * SPD + i] = s;
        }
}
template <int D, int SPD, class R>
SSDE_HD void blk_tvec(const R (&t)[SPD][SPD], const R* x, R* y) {
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
        for (int j = 0; j < SPD; ++j) {
            R s = 0.0;
#pragma unroll
            for (int i = 0; i < SPD; ++i) s += t[i][j] * x[d * SPD + i];
            y[d * SPD + j] = s;
        }
}

// ---------------------------------------------------------------------------------------------
// types
// ---------------------------------------------------------------------------------------------
template <int N, class R>
struct DState {                // predicted state: mean + packed symmetric covariance
    static constexpr int NS = N * (N + 1) / 2;
    R a[N];
    R P[NS];
};
template <int N, class R>
struct DAdj {                  // adjoint of a predicted state (Pbar: symmetric full-matrix adjoint, packed)
    static constexpr int NS = N * (N + 1) / 2;
    R a[N];
    R P[NS];
};
template <int N, class R>
struct DFwdElem {
    static constexpr int NS = N * (N + 1) / 2;
    R A[N][N];
    R b[N];
    R C[NS];
    R eta[N];
    R J[NS];
    static constexpr int NDBL = (N * N + 2 * N + 2 * NS) * ScalarOf<R>::NDBL;
};
template <int N, class R>
struct DBwdElem {
    static constexpr int NS = N * (N + 1) / 2;
    R L[N][N];
    R z[N];
    R D[NS];
    static constexpr int NDBL = (N * N + N + NS) * ScalarOf<R>::NDBL;
};
// measurement covariance of one row; `par`: H = sigma_obs^2 I (a parameter) rather than data
template <int D, class R>
struct DObsCov {
    R v[D][D];
    bool par;
};
// forward intermediates of one row that its adjoint needs
template <int D, int SPD, class R>
struct DAux {
    static constexpr int N = D * SPD;
    R Fi[D][D];
    R w[D];
    R G[N][D];
    R af[N];
    R TPf[N][N];
};

// ---------------------------------------------------------------------------------------------
// one row of the sequential filter (prediction form)
// ---------------------------------------------------------------------------------------------
// Update part shared by the filter step and by fwd_append: given the predicted (a, P) and y:
// u, F^-1, det F, w, G, then a <- a + G u, P <- P - G (Z P).
template <int D, int SPD, class R>
SSDE_HD void dense_update(R* a, R (&P)[D * SPD][D * SPD], const double* y, const DObsCov<D, R>& H, R (&Fi)[D][D],
                          R* w, R (&G)[D * SPD][D], R* u, R& det, R& quad) {
    constexpr int N = D * SPD;
    R F[D][D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        u[d] = y[d] - a[d * SPD];                                   // u = y - Z a, nllk_ctcrw.hpp:221
#pragma unroll
        for (int e = 0; e < D; ++e) F[d][e] = P[d * SPD][e * SPD] + H.v[d][e];      // :223
    }
    sym_inv_det(F, Fi, det);
    quad = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        R s = 0.0;
#pragma unroll
        for (int e = 0; e < D; ++e) s += Fi[d][e] * u[e];
        w[d] = s;
        quad += u[d] * s;
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int d = 0; d < D; ++d) {
            R s = 0.0;
#pragma unroll
            for (int e = 0; e < D; ++e) s += P[i][e * SPD] * Fi[e][d];
            G[i][d] = s;
        }
    R Pf[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        R s = a[i];
#pragma unroll
        for (int d = 0; d < D; ++d) s += G[i][d] * u[d];
        a[i] = s;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            R p = P[i][j];
#pragma unroll
            for (int d = 0; d < D; ++d) p -= G[i][d] * P[d * SPD][j];
            Pf[i][j] = p;
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) P[i][j] = (i == j) ? Pf[i][i] : 0.5 * (Pf[i][j] + Pf[j][i]);
}

// Advances `s` (predicted state of row i) to the predicted state of row i + 1.  F_out = det F
// (1 if missing), quad = u'F^-1 u (0 if missing); the row's log-likelihood contribution is
// -(log det F + quad)/2 (nllk_ctcrw.hpp:231-234).
template <int D, int SPD, bool WITH_AUX, class R>
SSDE_HD void dense_fwd_step(DState<D * SPD, R>& s, const StepBlk<SPD, R>& k, const double* y, const R* mu, bool has_obs,
                            const DObsCov<D, R>& H, DAux<D, SPD, R>* aux, R& F_out, R& quad_out) {
    constexpr int N = D * SPD;
    R P[N][N], af[N];
    sym_unpack<N>(s.P, P);
#pragma unroll
    for (int i = 0; i < N; ++i) af[i] = s.a[i];
    F_out = 1.0;
    quad_out = 0.0;
    if (has_obs) {
        R Fi[D][D], w[D], G[N][D], u[D];
        dense_update<D, SPD>(af, P, y, H, Fi, w, G, u, F_out, quad_out);
        if (WITH_AUX) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                aux->w[d] = w[d];
#pragma unroll
                for (int e = 0; e < D; ++e) aux->Fi[d][e] = Fi[d][e];
#pragma unroll
                for (int i = 0; i < N; ++i) aux->G[i][d] = G[i][d];
            }
        }
    } else if (WITH_AUX) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            aux->w[d] = 0.0;
#pragma unroll
            for (int e = 0; e < D; ++e) aux->Fi[d][e] = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) aux->G[i][d] = 0.0;
        }
    }
    // predict: a+ = T af + B mu,  P+ = T Pf T' + Q   (:238-241 / :216-217)
    R TPf[N][N], Pn[N][N];
    blk_left<D, SPD>(k.t, P, TPf);
    blk_right_t<D, SPD>(TPf, k.t, Pn);
    if (WITH_AUX) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            aux->af[i] = af[i];
#pragma unroll
            for (int j = 0; j < N; ++j) aux->TPf[i][j] = TPf[i][j];
        }
    }
    blk_vec<D, SPD>(k.t, af, s.a);
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
        for (int i = 0; i < SPD; ++i) {
            s.a[d * SPD + i] += k.b[i] * mu[d];
#pragma unroll
            for (int j = 0; j < SPD; ++j) Pn[d * SPD + i][d * SPD + j] += k.q[i][j];
        }
    sym_pack<N>(Pn, s.P);
}

// ---------------------------------------------------------------------------------------------
// forward scan elements
// ---------------------------------------------------------------------------------------------
template <int N, class R>
SSDE_HD DFwdElem<N, R> dfwd_identity() {
    DFwdElem<N, R> E;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        E.b[i] = 0.0; E.eta[i] = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) E.A[i][j] = (i == j) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int i = 0; i < DFwdElem<N, R>::NS; ++i) { E.C[i] = 0.0; E.J[i] = 0.0; }
    return E;
}

// E <- (track-start element with state s0) o E
template <int N, class R>
SSDE_HD void dfwd_append_start(DFwdElem<N, R>& E, const DState<N, R>& s0) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        E.b[i] = s0.a[i];
#pragma unroll
        for (int j = 0; j < N; ++j) E.A[i][j] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < DFwdElem<N, R>::NS; ++i) E.C[i] = s0.P[i];
}

// E <- (element of one ordinary row) o E:  (b, C) advance by one Kalman step, A <- T (I - G Z) A,
// eta += (Z A)' w,  J += (Z A)' F^-1 (Z A)   (see ctcrw_math.cuh fwd_append; H^-1 is never needed).
template <int D, int SPD, class R>
SSDE_HD void dfwd_append(DFwdElem<D * SPD, R>& E, const StepBlk<SPD, R>& k, const double* y, const R* mu, bool has_obs,
                         const DObsCov<D, R>& H) {
    constexpr int N = D * SPD;
    R C[N][N];
    sym_unpack<N>(E.C, C);
    if (has_obs) {
        R Fi[D][D], w[D], G[N][D], u[D], det, quad;
        R J[N][N];
        sym_unpack<N>(E.J, J);
        dense_update<D, SPD>(E.b, C, y, H, Fi, w, G, u, det, quad);
        R ZA[D][N], FZA[D][N];
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
            for (int c = 0; c < N; ++c) ZA[d][c] = E.A[d * SPD][c];
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
            for (int c = 0; c < N; ++c) {
                R s = 0.0;
#pragma unroll
                for (int e = 0; e < D; ++e) s += Fi[d][e] * ZA[e][c];
                FZA[d][c] = s;
            }
#pragma unroll
        for (int c = 0; c < N; ++c) {
#pragma unroll
            for (int d = 0; d < D; ++d) E.eta[c] += ZA[d][c] * w[d];
#pragma unroll
            for (int c2 = 0; c2 < N; ++c2)
#pragma unroll
                for (int d = 0; d < D; ++d) J[c][c2] += ZA[d][c] * FZA[d][c2];
        }
        sym_pack<N>(J, E.J);
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int c = 0; c < N; ++c)
#pragma unroll
                for (int d = 0; d < D; ++d) E.A[i][c] -= G[i][d] * ZA[d][c];
    }
    R TA[N][N], TC[N][N], Cn[N][N], tb[N];
    blk_left<D, SPD>(k.t, E.A, TA);
    blk_left<D, SPD>(k.t, C, TC);
    blk_right_t<D, SPD>(TC, k.t, Cn);
    blk_vec<D, SPD>(k.t, E.b, tb);
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) E.A[i][j] = TA[i][j];
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
        for (int i = 0; i < SPD; ++i) {
            E.b[d * SPD + i] = tb[d * SPD + i] + k.b[i] * mu[d];
#pragma unroll
            for (int j = 0; j < SPD; ++j) Cn[d * SPD + i][d * SPD + j] += k.q[i][j];
        }
    sym_pack<N>(Cn, E.C);
}

// General composition (`Ei` earlier rows, `Ej` later rows), formulas in ctcrw_math.cuh fwd_combine.
template <int N, class R>
SSDE_HD DFwdElem<N, R> dfwd_combine(const DFwdElem<N, R>& Ei, const DFwdElem<N, R>& Ej) {
    DFwdElem<N, R> Ro;
    R C[N][N], J[N][N], X[N][N], M[N][N], AM[N][N], T1[N][N], T2[N][N];
    sym_unpack<N>(Ei.C, C);
    sym_unpack<N>(Ej.J, J);
    mat_mul<N>(C, J, X);
#pragma unroll
    for (int i = 0; i < N; ++i) X[i][i] += 1.0;
    inv_general<N>(X, M);
    mat_mul<N>(Ej.A, M, AM);
    mat_mul<N>(AM, Ei.A, Ro.A);
    // C = A_j M C_i A_j' + C_j
    mat_mul<N>(AM, C, T1);
    mat_mul_nt<N>(T1, Ej.A, T2);
    sym_pack<N>(T2, Ro.C);
#pragma unroll
    for (int i = 0; i < DFwdElem<N, R>::NS; ++i) Ro.C[i] += Ej.C[i];
    // J = A_i' M' J_j A_i + J_i
    mat_mul_tn<N>(M, J, T1);               // M' J_j
    mat_mul<N>(T1, Ei.A, T2);              // M' J_j A_i
    mat_mul_tn<N>(Ei.A, T2, T1);           // A_i' M' J_j A_i
    sym_pack<N>(T1, Ro.J);
#pragma unroll
    for (int i = 0; i < DFwdElem<N, R>::NS; ++i) Ro.J[i] += Ei.J[i];
    // b = A_j M (b_i + C_i eta_j) + b_j ;  eta = A_i' M' (eta_j - J_j b_i) + eta_i
    R t[N], r[N], v[N];
    mat_vec<N>(C, Ej.eta, t);
    mat_vec<N>(J, Ei.b, r);
#pragma unroll
    for (int i = 0; i < N; ++i) { t[i] += Ei.b[i]; r[i] = Ej.eta[i] - r[i]; }
    mat_vec<N>(AM, t, Ro.b);
    mat_tvec<N>(M, r, v);
    mat_tvec<N>(Ei.A, v, Ro.eta);
#pragma unroll
    for (int i = 0; i < N; ++i) { Ro.b[i] += Ej.b[i]; Ro.eta[i] += Ei.eta[i]; }
    return Ro;
}

// State after pushing `s` through element E
template <int N, class R>
SSDE_HD DState<N, R> dfwd_apply(const DFwdElem<N, R>& E, const DState<N, R>& s) {
    DState<N, R> ro;
    R P[N][N], J[N][N], X[N][N], M[N][N], AM[N][N], T1[N][N], T2[N][N];
    sym_unpack<N>(s.P, P);
    sym_unpack<N>(E.J, J);
    mat_mul<N>(P, J, X);
#pragma unroll
    for (int i = 0; i < N; ++i) X[i][i] += 1.0;
    inv_general<N>(X, M);
    mat_mul<N>(E.A, M, AM);
    mat_mul<N>(AM, P, T1);
    mat_mul_nt<N>(T1, E.A, T2);
    sym_pack<N>(T2, ro.P);
#pragma unroll
    for (int i = 0; i < DState<N, R>::NS; ++i) ro.P[i] += E.C[i];
    R t[N];
    mat_vec<N>(P, E.eta, t);
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] += s.a[i];
    mat_vec<N>(AM, t, ro.a);
#pragma unroll
    for (int i = 0; i < N; ++i) ro.a[i] += E.b[i];
    return ro;
}

// ---------------------------------------------------------------------------------------------
// adjoint (reverse-time) elements
// ---------------------------------------------------------------------------------------------
template <int N, class R>
SSDE_HD DAdj<N, R> dadj_zero() {
    DAdj<N, R> g;
#pragma unroll
    for (int i = 0; i < N; ++i) g.a[i] = 0.0;
#pragma unroll
    for (int i = 0; i < DAdj<N, R>::NS; ++i) g.P[i] = 0.0;
    return g;
}
template <int N, class R>
SSDE_HD DBwdElem<N, R> dbwd_identity() {
    DBwdElem<N, R> E;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        E.z[i] = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) E.L[i][j] = (i == j) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int i = 0; i < DBwdElem<N, R>::NS; ++i) E.D[i] = 0.0;
    return E;
}
template <int N, class R>
SSDE_HD DBwdElem<N, R> dbwd_const(const DAdj<N, R>& g) {
    DBwdElem<N, R> E;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        E.z[i] = -g.a[i];
#pragma unroll
        for (int j = 0; j < N; ++j) E.L[i][j] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < DBwdElem<N, R>::NS; ++i) E.D[i] = g.P[i];
    return E;
}

// Elementary backward element of one ordinary row (see the header comment); `cut`: last row of
// its track (the incoming adjoint is discarded, L = 0).
template <int D, int SPD, class R>
SSDE_HD DBwdElem<D * SPD, R> dbwd_row_elem(const StepBlk<SPD, R>& k, const DAux<D, SPD, R>& ax, bool has_obs, bool cut) {
    constexpr int N = D * SPD;
    DBwdElem<N, R> E;
    R Dm[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        E.z[i] = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) { Dm[i][j] = 0.0; E.L[i][j] = 0.0; }
    }
    if (has_obs) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            E.z[d * SPD] = ax.w[d];
#pragma unroll
            for (int e = 0; e < D; ++e) Dm[d * SPD][e * SPD] = 0.5 * (ax.Fi[d][e] - ax.w[d] * ax.w[e]);
        }
    }
    sym_pack<N>(Dm, E.D);
    if (!cut) {
        // L = T (I - G Z):  start from I - G Z, multiply by the blocks
        R X[N][N];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) X[i][j] = (i == j) ? 1.0 : 0.0;
        if (has_obs) {
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int d = 0; d < D; ++d) X[i][d * SPD] -= ax.G[i][d];
        }
        blk_left<D, SPD>(k.t, X, E.L);
    }
    return E;
}

// Composition: E1 covers EARLIER rows, E2 later rows.
//   L = L2 L1;  z = L1' z2 + z1;  D = L1' D2 L1 + D1 - sym(L1' z2 z1')
template <int N, class R>
SSDE_HD DBwdElem<N, R> dbwd_combine(const DBwdElem<N, R>& E1, const DBwdElem<N, R>& E2) {
    DBwdElem<N, R> Ro;
    R D2[N][N], T1[N][N], T2[N][N], t[N];
    mat_mul<N>(E2.L, E1.L, Ro.L);
    sym_unpack<N>(E2.D, D2);
    mat_mul<N>(D2, E1.L, T1);
    mat_mul_tn<N>(E1.L, T1, T2);
    mat_tvec<N>(E1.L, E2.z, t);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        Ro.z[i] = t[i] + E1.z[i];
#pragma unroll
        for (int j = 0; j < N; ++j) T2[i][j] -= t[i] * E1.z[j];
    }
    sym_pack<N>(T2, Ro.D);
#pragma unroll
    for (int i = 0; i < DBwdElem<N, R>::NS; ++i) Ro.D[i] += E1.D[i];
    return Ro;
}

// Adjoint at the start of E's range given the adjoint `g` flowing in at its end.
template <int N, class R>
SSDE_HD DAdj<N, R> dbwd_apply(const DBwdElem<N, R>& E, const DAdj<N, R>& g) {
    DAdj<N, R> ro;
    R P[N][N], T1[N][N], T2[N][N], t[N];
    sym_unpack<N>(g.P, P);
    mat_mul<N>(P, E.L, T1);
    mat_mul_tn<N>(E.L, T1, T2);
    mat_tvec<N>(E.L, g.a, t);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        ro.a[i] = t[i] - E.z[i];
#pragma unroll
        for (int j = 0; j < N; ++j) T2[i][j] += t[i] * E.z[j];
    }
    sym_pack<N>(T2, ro.P);
#pragma unroll
    for (int i = 0; i < DAdj<N, R>::NS; ++i) ro.P[i] += E.D[i];
    return ro;
}

// Adjoints of the step blocks of one row given the adjoint `g` of the state the row predicts:
// bar.t, bar.q, bar.b (each entry of a block treated as an independent variable, summed over the
// dimensions), gmu[d] = d nllk / d mu_d, and g_h = d nllk / d h when H = h I (else 0).
template <int D, int SPD, class R>
SSDE_HD void dense_step_adjoint(const DAdj<D * SPD, R>& g, const StepBlk<SPD, R>& k, const DAux<D, SPD, R>& ax, const R* mu,
                                bool has_obs, bool h_par, StepBlk<SPD, R>& bar, R* gmu, R& g_h) {
    constexpr int N = D * SPD;
    R Pb[N][N];
    sym_unpack<N>(g.P, Pb);
#pragma unroll
    for (int i = 0; i < SPD; ++i) {
        bar.b[i] = 0.0;
#pragma unroll
        for (int j = 0; j < SPD; ++j) { bar.t[i][j] = 0.0; bar.q[i][j] = 0.0; }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        gmu[d] = 0.0;
#pragma unroll
        for (int i = 0; i < SPD; ++i) {
            gmu[d] += k.b[i] * g.a[d * SPD + i];
            bar.b[i] += g.a[d * SPD + i] * mu[d];
#pragma unroll
            for (int j = 0; j < SPD; ++j) {
                R s = g.a[d * SPD + i] * ax.af[d * SPD + j];
#pragma unroll
                for (int m = 0; m < N; ++m) s += 2.0 * (Pb[d * SPD + i][m] * ax.TPf[m][d * SPD + j]);
                bar.t[i][j] += s;
                bar.q[i][j] += Pb[d * SPD + i][d * SPD + j];
            }
        }
    }
    g_h = 0.0;
    if (has_obs && h_par) {
        R afb[N], T1[N][N], Pfb[N][N];
        blk_tvec<D, SPD>(k.t, g.a, afb);                // af_bar = T' abar+
        blk_left_t<D, SPD>(k.t, Pb, T1);                // Pf_bar = T' Pbar+ T
        blk_right<D, SPD>(T1, k.t, Pfb);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            R ga = 0.0, gpg = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                ga += ax.G[i][d] * afb[i];
                R s = 0.0;
#pragma unroll
                for (int j = 0; j < N; ++j) s += Pfb[i][j] * ax.G[j][d];
                gpg += ax.G[i][d] * s;
            }
            g_h += 0.5 * (ax.Fi[d][d] - ax.w[d] * ax.w[d]) - ga * ax.w[d] + gpg;
        }
    }
}

}  // namespace ssde
