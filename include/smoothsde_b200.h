/* smoothsde_b200 -- C ABI of the B200-native likelihood engine for smoothSDE's hot path.
 *
 * Drop-in boundary.  In the reference, R reaches the objective through TMB's registered .Call
 * routines (src/init.c:6-15): MakeADFunObject(data, parameters, reportenv, control) deep-copies
 * the data list into a C++ objective_function and tapes src/smoothSDE.cpp:9-28; EvalADFunObject
 * (ptr, theta, control) replays the tape for value / gradient; REPORT(aest_all)
 * (src/nllk/nllk_ctcrw.hpp:249) is read back through the report environment.  This header is
 * what an R (or any other) host binds instead:
 *
 *   reference (.Call, src/init.c)      this library
 *   -------------------------------    ------------------------------------------------------
 *   MakeADFunObject        :6          ssde_create        (copies the data list to the GPU once)
 *   EvalADFunObject        :8          ssde_eval          (order 0: nllk, order 1: + gradient)
 *   MakeADGradObject       :12         ssde_eval(order = 1)
 *   MakeADHessObject2      :13         ssde_eval(order = 2), ssde_hvp, ssde_hess_cols_device
 *   REPORT(aest_all)                   ssde_report
 *   getParameterOrder      :11         ssde_n_par / ssde_par_layout
 *   (external pointer finalizer)       ssde_destroy
 *   Rf_error                           non-zero return + ssde_last_error / ssde_create_error
 *
 * All entry points are extern "C", take plain pointers and sizes, never throw and never call
 * back into the host.  Everything is fp64.  A handle must not be used concurrently from two
 * threads.  There is no CPU fallback: every call fails with SSDE_ERR_CUDA if no device exists.
 */
#ifndef SMOOTHSDE_B200_H
#define SMOOTHSDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssde_handle ssde_handle;

/* DATA_STRING(type) of src/smoothSDE.cpp:11-25.  BM_t, CIR, ESEAL_SSM are not built (the host
 * layer answers SSDE_ERR_UNSUPPORTED; the reference says error("Unknown SDE type") only for names
 * outside its list, smoothSDE.cpp:25 -> SSDE_ERR_UNKNOWN_TYPE here).
 * BM_SSM / OU_SSM (nllk_bm_ssm.hpp, nllk_ou_ssm.hpp): a0 is [n_ID x n_dim] (first observation of
 * each track, R/sde.R:549-550), P0 [n_dim x n_dim] (default diag(10), R/sde.R:553); parameter
 * vector as for CTCRW.
 * Kalman models: with H = sigma_obs^2 I and a P0 of the default shape (CTCRW: block-diagonal with
 * identical 2x2 blocks; SSM: c I) the dimensions decouple and the fast kernels run; a user H_array
 * (nllk_ctcrw.hpp:203-205) or any other symmetric P0 selects the coupled filter (full N x N
 * covariance recursion, same scan kernels with matrix-valued elements). */
enum ssde_model { SSDE_BM = 0, SSDE_OU = 1, SSDE_CTCRW = 2, SSDE_BM_SSM = 3, SSDE_OU_SSM = 4 };

enum ssde_status {
    SSDE_OK = 0,
    SSDE_ERR_UNKNOWN_TYPE = 1,
    SSDE_ERR_BAD_ARG = 2,
    SSDE_ERR_UNSUPPORTED = 3,
    SSDE_ERR_CUDA = 4,
    SSDE_ERR_NUMERIC = 5      /* device status word: innovation variance F <= 0 in the filter (bit 1), scan
                               * look-back time-out (bit 0); Laplace: H_bb not positive definite, ... */
};

/* dgTMatrix triplets as produced by as_sparse(), R/utility.R:204-213: 0-based i, j; duplicate
 * (i, j) entries are summed. */
typedef struct {
    int64_t nrow, ncol, nnz;
    const int32_t* i;
    const int32_t* j;
    const double* x;
} ssde_triplet;

enum ssde_shard_flags {
    SSDE_SHARD_CONT_PREV = 1,  /* first row continues a track that started on the previous shard */
    SSDE_SHARD_CONT_NEXT = 2,  /* the track of the last row continues on the next shard */
    SSDE_SHARD_NO_PENALTY = 4  /* smoothing penalty is added by another shard (rank != 0) */
};

/* The data list that SDE$setup() hands to MakeADFun (R/sde.R:528-536, :569-598), for the rows
 * of one shard (a whole problem is one shard with shard_flags = 0). */
typedef struct {
    int32_t model;            /* enum ssde_model */
    int32_t n_dim;            /* ncol(obs) = length(response) */
    int64_t n;                /* rows */
    const double* ID;         /* [n]  factor codes; only equality of neighbours is used */
    const double* times;      /* [n] */
    const double* obs;        /* [n x n_dim] column-major (R matrix); NaN/NA = missing */
    ssde_triplet X_fe;        /* [(n_par*n) x p_fe]; rows j*n..(j+1)*n-1 belong to parameter j */
    ssde_triplet X_re;        /* [(n_par*n) x p_re] */
    ssde_triplet S;           /* [p_re x p_re] block-diagonal penalty */
    int32_t n_smooth;         /* length(ncol_re) (1 with ncol_re[0] = 0 when there are no smooths) */
    const int32_t* ncol_re;   /* [n_smooth] */
    int32_t include_penalty;  /* only honoured for BM/OU, as in the reference (nllk_sde.hpp:91) */
    int32_t n_ID;             /* Kalman models: rows of a0 (tracks starting on this shard) */
    const double* a0;         /* CTCRW: [n_ID x 2*n_dim] column-major, (x, 0, y, 0, ...) R/sde.R:574-580;
                                 BM_SSM / OU_SSM: [n_ID x n_dim] */
    const double* P0;         /* CTCRW: [2*n_dim x 2*n_dim] column-major, R/sde.R:582-588; SSM: [n_dim x n_dim] */
    const double* H_array;    /* Kalman models: NULL or array(0) (H_len <= 1) for H = sigma_obs^2 I, else the user's
                                 measurement covariances [n_dim x n_dim x n] column-major (R/sde.R:593-598); the
                                 objective then does not depend on log_sigma_obs (its gradient entry is 0) */
    int64_t H_len;            /* length(H_array): n_dim * n_dim * n, or <= 1 */
    int32_t device;           /* CUDA device ordinal */
    int32_t shard_flags;      /* enum ssde_shard_flags */
    double t_next;            /* time of the next shard's first row (only with CONT_NEXT) */
    /* decay terms of nllk_sde (BM / OU), nllk_sde.hpp:31-33,47-59, R/sde.R:162-177,636-648: column
     * col_decay[i] (1-based) of X_re is multiplied, row by row, by exp(-exp(log_decay[ind_decay[i]]) * t_decay).
     * t_decay = NULL or t_decay_len <= 1: no decay term (R passes t_decay = 0 and maps log_decay off). */
    const double* t_decay;    /* [n_par * n], one entry per row of X_re */
    int64_t t_decay_len;
    const int32_t* col_decay; /* [n_col_decay] */
    const int32_t* ind_decay; /* [n_col_decay], values 1 .. n_decay (n_decay <= 16) */
    int32_t n_col_decay;
} ssde_desc;

/* Zero-copy variant for data that already lives on the GPU in the engine's native "warp-tile
 * transposed" layout (smoothsde_b200/csrc/design.cuh): rows are grouped in warp-tiles of
 * WT = 32 * LC rows, row i = q*WT + l*LC + k lives at position q*WT + k*32 + l of every per-row
 * array, and the design is a sliced-ELL structure with one 24-byte descriptor per warp-tile.
 * ssde_layout_info reports {LC, WT, padding unit, sizeof(descriptor)}; ssde_padded_rows(n) is the
 * length every per-row array must have.  All d_* pointers are device pointers that must stay
 * valid for the life of the handle; the rest is host memory and is copied. */
typedef struct {
    int64_t val_off;          /* first value of the warp-tile in d_val */
    int64_t col_off;          /* first column index of the warp-tile in d_col */
    uint32_t kmax;            /* byte p = slots of SDE parameter p; S = sum of the bytes */
    uint32_t flags;           /* bit0: one column list col[col_off + j] for the whole warp-tile,
                                 else per nonzero, indexed like the values.
                                 bits 8+2p, 9+2p (bit0 set only): 0 = parameter p stores its own values;
                                 t+1 = parameter p is an ALIAS of parameter t < p -- its kmax_p value slots
                                 hold the same numbers as t's in every row (tau ~ s(time), nu ~ s(time):
                                 one basis, two column blocks) and are not stored.  It keeps its own
                                 column slots.  Not accepted together with decay terms. */
} ssde_wt_desc;               /* value of slot j, row (k, l): d_val[val_off + (k*SV + j)*32 + l], SV = the
                                 kmax bytes of the non-aliased parameters summed (= S without aliases),
                                 value slots numbered over the non-aliased parameters in order */

typedef struct {
    int32_t model, n_dim, n_par;
    int64_t n, n_pad;
    int64_t nnz;              /* informational */
    const ssde_wt_desc* d_desc; /* [n_pad / WT] */
    const double* d_val;
    const uint32_t* d_col;    /* columns in theta = [coeff_fe | coeff_re] */
    const double* d_obs;      /* n_dim planes of n_pad doubles, permuted; missing entries hold any finite value */
    const double* d_dt;       /* [n_pad] permuted; dt_i = t_{i+1} - t_i; 1 on rows flagged LAST; CTCRW: on
                                 track-start rows the 0-based track index (row of a0) instead */
    const uint8_t* d_flags;   /* [n_pad] permuted; bit0 track start, bit1 track last, bit2 obs present
                                 (CTCRW), bit(3+d) dimension d missing (BM/OU); 0xff = beyond row n */
    int32_t p_fe, p_re;
    ssde_triplet S;
    int32_t n_smooth;
    const int32_t* ncol_re;
    int32_t include_penalty;
    int32_t n_ID;
    const int64_t* track_starts; /* [n_ID] host, sorted rows flagged as track start */
    const double* a0;            /* [n_ID x 2*n_dim] host, ROW-major */
    double P0[3];                /* shared 2x2 block (p11, p12, p22) */
    int32_t device;
    int32_t shard_flags;
    const int32_t* mu_cols;      /* CTCRW, optional (host): every column of theta that some mu_d predictor uses; */
    int32_t n_mu_cols;           /* lets the kernels skip B*mu when all those entries are exactly 0.  NULL: unknown */
} ssde_packed_desc;

int64_t ssde_padded_rows(int64_t n);
int ssde_layout_info(int32_t info[4]);

/* Host-side packing of the design of `desc` into the device layout (no GPU needed); free the
 * arrays with ssde_pack_free. */
typedef struct {
    int64_t n_pad, n_desc, n_val, n_col;
    ssde_wt_desc* desc;
    double* val;
    uint32_t* col;
} ssde_host_pack;
int ssde_pack_host(const ssde_desc* desc, ssde_host_pack* out);
void ssde_pack_free(ssde_host_pack* p);

int ssde_create(const ssde_desc* desc, ssde_handle** out);
int ssde_create_packed(const ssde_packed_desc* desc, ssde_handle** out);
void ssde_destroy(ssde_handle* h);

/* Length of the joint parameter vector and its layout (PARAMETER order of the templates):
 *   CTCRW : log_sigma_obs, coeff_fe[p_fe], log_lambda[n_smooth], coeff_re[p_re]
 *           (nllk_ctcrw.hpp:135-140)
 *   BM/OU : coeff_fe[p_fe], log_lambda[n_smooth], log_decay[n_decay], coeff_re[p_re]
 *           (nllk_sde.hpp:42-45; without decay terms log_decay is mapped off, R/sde.R:648, and
 *           n_decay = 0)
 * offsets[4] = {log_sigma_obs (or -1), coeff_fe, log_lambda, coeff_re};
 * ssde_decay_layout: offset (or -1) and length of log_decay. */
int ssde_n_par(const ssde_handle* h);
int ssde_par_layout(const ssde_handle* h, int32_t offsets[4], int32_t sizes[4]);
int ssde_decay_layout(const ssde_handle* h, int32_t* offset, int32_t* size);

/* One objective evaluation with HOST buffers (what obj$fn / obj$gr / obj$he do through
 * EvalADFunObject and the MakeADHessObject2 tape).
 *   order 0: *nllk;  order 1: *nllk and grad[ssde_n_par];
 *   order 2: also hess[ssde_n_par x ssde_n_par] (column-major, symmetric): the exact joint Hessian
 *            of the penalised objective, one tangent pass per column (see ssde_hvp).
 * Copies `par` to the device, runs the kernels, copies the result back, synchronises. */
int ssde_eval(ssde_handle* h, const double* par, int order, double* nllk, double* grad, double* hess);

/* Same evaluation with DEVICE buffers and no synchronisation: d_out[0] = nllk (this shard's
 * part), d_out[1 .. n_par] = gradient.  `stream` is a cudaStream_t (NULL = the handle's own
 * stream).  Used by the multi-GPU host, which all-reduces d_out over NCCL. */
int ssde_eval_device(ssde_handle* h, const double* d_par, int order, double* d_out, void* stream);
/* Exact Hessian-vector products (second-order adjoint).  The kernels are templates over their
 * scalar type; instantiated with (value, tangent) pairs, one forward + adjoint pass along a
 * direction v of the parameter vector returns nllk, the gradient and H v, H the Hessian of the
 * joint penalised objective -- what TMB obtains by taping the gradient tape again
 * (MakeADHessObject2, src/init.c:13) and needs for the Laplace approximation over coeff_re
 * (R/sde.R:522-524) and for obj$he (R/sde.R:1363).
 *   ssde_hvp: HOST buffers; dirs = [n_par x n_dir] column-major directions; hv likewise = H dirs;
 *             nllk / grad may be NULL.
 *   ssde_hvp_device: DEVICE buffers, asynchronous on `stream`; d_out[1 + n_par + 1] as in
 *             ssde_eval_device, d_hv[n_par] = H d_dir (this shard's part: sum over shards).
 *   ssde_hess_cols_device: columns first .. first+count-1 of H into d_hess[n_par x count]
 *             (column-major), e.g. the coeff_re block for the Laplace inner problem.
 * Not available for time shards (SSDE_SHARD_CONT_*) yet: SSDE_ERR_UNSUPPORTED. */
int ssde_hvp(ssde_handle* h, const double* par, int n_dir, const double* dirs, double* nllk, double* grad, double* hv);
int ssde_hvp_device(ssde_handle* h, const double* d_par, const double* d_dir, double* d_out, double* d_hv, void* stream);
int ssde_hess_cols_device(ssde_handle* h, const double* d_par, int first, int count, double* d_out, double* d_hess, void* stream);
/* BM / OU without decay terms (designs whose warp-tiles are uniform, <= 32 slots, <= 12 per SDE parameter):
 * the Hessian of the penalised objective with respect to theta = [coeff_fe | coeff_re] in ONE pass over the
 * design, H = X' W X + blockdiag(0, lambda_i S_i), with the exact NP x NP second-derivative block W_i of
 * every row (rows are independent, nllk_sde.hpp:73-84) -- the "X_re' W X_re + S_lambda" of the Laplace inner
 * problem.  Tangent passes (ssde_hess_cols_device) need one sweep per column, i.e. thousands for
 * s(ID, bs = "re") with one level per track.  d_hess: [p_theta x p_theta] column-major DEVICE buffer,
 * overwritten (this shard's part: sum over shards; the penalty is added by the shard that owns it).
 * Asynchronous on `stream`.  SSDE_ERR_UNSUPPORTED for the Kalman models and decay models. */
int ssde_hess_theta_device(ssde_handle* h, const double* d_par, double* d_hess, void* stream);

/* Laplace-marginal objective over coeff_re: what MakeADFun(..., random = "coeff_re") evaluates
 * (R/sde.R:522-524, :656-658; TMB's inner newton() + sparse Cholesky):
 *     f(theta) = g(theta, b_hat) + 1/2 log det H_bb(theta, b_hat) - n_b/2 log(2 pi),
 *     b_hat = argmin_b g(theta, b),   g = the joint penalised nllk.
 * The inner Newton iterations use the exact H_bb (tangent passes) and cuSOLVER's dense Cholesky
 * (potrf / potrs) on the device; the gradient of f costs 2 n_b (4 n_b with Richardson
 * extrapolation) further tangent passes, see csrc/ssde_laplace.cu.
 *   ssde_laplace_eval: `par` [ssde_n_par] holds the outer parameters and the STARTING value of
 *     coeff_re on entry, and coeff_re = b_hat on exit (TMB's env$last.par).  order 0: *value;
 *     order 1: also grad[ssde_n_par], the gradient of f w.r.t. every non-random entry (0 in the
 *     coeff_re slots).  The host adapter applies `map` on top, as for ssde_eval. */
typedef struct ssde_laplace ssde_laplace;
typedef struct {
    int32_t max_newton;       /* inner Newton iterations (default 100) */
    int32_t richardson;       /* 1 (default): O(eps^4) differences of the Hessian-vector products */
    double grad_tol;          /* inner convergence: max |d g / d b| (default 1e-8) */
    double fd_step;           /* eps of those differences, along unit directions (default 1e-3) */
} ssde_laplace_opts;
typedef struct {
    double joint;             /* g(theta, b_hat) */
    double logdet;            /* log det H_bb(theta, b_hat) */
    double grad_max;          /* max |d g / d b| at exit */
    int32_t converged, n_newton, n_hess, n_value, n_hvp;
} ssde_laplace_info;
int ssde_laplace_create(ssde_handle* h, const ssde_laplace_opts* opts, ssde_laplace** out);
void ssde_laplace_destroy(ssde_laplace* w);
int ssde_laplace_eval(ssde_laplace* w, double* par, int order, double* value, double* grad, ssde_laplace_info* info);
/* H_bb [n_b x n_b, column-major] at the mode of the last ssde_laplace_eval */
int ssde_laplace_hessian_bb(ssde_laplace* w, double* hess_bb);
const char* ssde_laplace_error(const ssde_laplace* w);

/* Diagnostics (libraries built with -DSSDE_STATS only, else SSDE_ERR_UNSUPPORTED): look-back and
 * per-phase cycle counters of the scan kernels, [0..15] forward, [16..31] adjoint. */
int ssde_debug_stats(ssde_handle* h, uint64_t out[32], int reset);
/* Diagnostics: the look-back of the scan kernels stops at an aggregate whose linear part has decayed
 * below `tol` (default 1e-60: the filter has forgotten its initial condition, csrc/models.cuh).
 * tol < 0 switches that shortcut off for every handle on `device`, so tests can compare both paths. */
int ssde_debug_const_map_tol(int device, double tol);

/* The handle's CUDA device ordinal and its own stream (a cudaStream_t). */
int ssde_device(const ssde_handle* h);
void* ssde_stream(const ssde_handle* h);

/* Time-sharded evaluation of ONE long CTCRW track whose rows are split along time over several
 * handles / ranks (shard_flags SSDE_SHARD_CONT_PREV / CONT_NEXT).  The filter and its adjoint are
 * associative scans, so each shard is summarised by one composite element; the host gathers the
 * elements of all shards between the stages (two all-gathers of < 1 KB and the final all-reduce
 * per evaluation).  Everything is asynchronous on `stream`, all pointers are device pointers:
 *   stage 0: forward summary pass over the TAIL of the shard (its last 4096 rows);
 *            d_out[ssde_shard_elem_doubles(h, 0)] = composite element + constant-map flag (last double)
 *   stage 3: the same over the WHOLE shard -- needed only when some shard's flag of stage 0 is 0,
 *            i.e. its tail does not make the filter forget what came before (long observation gaps)
 *   stage 1: d_elems = forward elements of all n_shards shards (shard-major); computes the incoming
 *            state, runs the forward pass and the adjoint summary pass over the HEAD of the shard;
 *            d_out[ssde_shard_elem_doubles(h, 1)] = this shard's adjoint element + flag
 *   stage 4: the adjoint summary over the whole shard (fallback as stage 3)
 *   stage 2: d_elems = adjoint elements of all shards; computes the incoming adjoint, runs the
 *            adjoint pass;  d_out[1 + n_par + 1] = partial nllk, gradient, status (as ssde_eval_device)
 */
int ssde_shard_elem_doubles(const ssde_handle* h, int which);
int ssde_eval_stage(ssde_handle* h, const double* d_par, int stage, const double* d_elems, int n_shards,
                    int my_shard, double* d_out, void* stream);

/* Check the device-side status word of the last evaluation (synchronises the stream). */
int ssde_check(ssde_handle* h);

/* REPORT(aest_all): [n x 2*n_dim] column-major predicted-state means after each row
 * (nllk_ctcrw.hpp:192-194, :246-249) at the parameters of the last ssde_eval.
 * One documented difference: the row that ENDS a track holds, in the reference, a prediction made
 * with the cross-track dt = times[first row of next track] - times[last row] (:126-129, :206-208) --
 * a value the loop discards at the ID change (:196-200) and that overflows when the next track's
 * clock restarts; here that row is predicted with dt = 1 (the value the reference uses for the very
 * last row, :129).  Every other row agrees with the reference to 1e-9 (tests/test_gpu_ref.py). */
int ssde_report(ssde_handle* h, double* aest_all);

/* Device time in milliseconds of the kernels of the last ssde_eval / ssde_eval_device on the
 * handle's own stream (CUDA events around the launches). */
double ssde_last_eval_ms(ssde_handle* h);
/* Number of kernels the last evaluation launched. */
int ssde_last_eval_launches(const ssde_handle* h);

/* Launch geometry of the handle's kernels: {SMs, forward grid, adjoint grid, BM/OU grid, forward
 * tiles, adjoint tiles, BM/OU tiles, tracks}.  Grids are (resident CTAs per SM) x SMs. */
int ssde_launch_info(const ssde_handle* h, int32_t info[8]);

/* Per-kernel device times: with profiling on, every evaluation records a CUDA event in front of
 * each kernel on the launching stream; ssde_last_kernel_times returns how many kernels the last
 * evaluation ran and fills ms[] / names[] (static strings) for up to `cap` of them. */
int ssde_set_profile(ssde_handle* h, int on);
int ssde_last_kernel_times(ssde_handle* h, int cap, float* ms, const char** names);

/* Exact-transition CTCRW simulator on the device (SDE$simulate, R/sde.R:1448-1478; CTCRW_cov,
 * R/utility.R:188-196), one dimension: all arrays are [n_tracks x n_steps] track-major device
 * arrays; d_z[track][0] holds the start position on entry; d_e1, d_e2 are standard normal draws;
 * d_mu may be NULL (mu = 0).  Asynchronous on `stream`. */
int ssde_simulate_ctcrw(int device, int64_t n_tracks, int64_t n_steps, const double* d_times,
                        const double* d_tau, const double* d_nu, const double* d_mu,
                        const double* d_e1, const double* d_e2, double* d_z, void* stream);
/* Exact-transition OU simulator (R/sde.R:1439-1447), same conventions: z_i | z_{i-1} uses the
 * natural-scale parameters mu, tau, kappa of row i-1; d_e are standard normal draws. */
int ssde_simulate_ou(int device, int64_t n_tracks, int64_t n_steps, const double* d_times, const double* d_mu,
                     const double* d_tau, const double* d_kappa, const double* d_e, double* d_z, void* stream);

const char* ssde_last_error(const ssde_handle* h);
const char* ssde_create_error(void);
const char* ssde_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SMOOTHSDE_B200_H */
