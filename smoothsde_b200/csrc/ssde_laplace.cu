// Laplace-marginal objective over the random effects coeff_re, on the device.
//
// In the reference, SDE$setup() asks TMB for MakeADFun(..., random = "coeff_re")
// (R/sde.R:522-524, :656-658); every obj$fn(theta) then runs TMB's inner Newton optimisation
// over coeff_re with the AD-of-AD sparse Hessian (MakeADHessObject2, src/init.c:13) and a
// CHOLMOD factorisation, and returns
//     f(theta) = g(theta, b_hat) + 1/2 log det H_bb(theta, b_hat) - n_b/2 log(2 pi),
// g = the joint penalised nllk of src/nllk/*.hpp.  Here:
//   * g, its gradient and the EXACT columns of its Hessian come from the engine's kernels
//     (ssde_eval_device / ssde_hess_cols_device / ssde_hvp_device: tangent passes, dual.cuh);
//   * the Newton system and the log-determinant use cuSOLVER's dense Cholesky (potrf / potrs) of
//     the n_b x n_b block -- the block is small (tens to hundreds) and dense for spline smooths;
//   * the gradient of f needs third derivatives of g through log det H_bb.  With H_bb^-1 =
//     Z Z' (Z = L^-T),  d/dx_k log det H_bb = sum_j D^3 g[z_j, z_j, e_k]; each term is a central
//     difference (Richardson-extrapolated) of Hessian-vector products along z_j, so ALL
//     components cost 4 n_b tangent passes, independent of the number of outer parameters.  The
//     implicit dependence of b_hat on theta is eliminated with the same Cholesky factor:
//         grad f = g_theta + 1/2 (w_theta - H_theta,b H_bb^-1 w_b).
// Everything is written on top of the public C ABI; only small vectors cross PCIe.
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <cmath>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../include/smoothsde_b200.h"

struct ssde_laplace {
    ssde_handle* h = nullptr;
    int device = 0, np = 0, o_re = 0, nb = 0;
    cudaStream_t st = nullptr;
    cusolverDnHandle_t cs = nullptr;
    double *d_par = nullptr, *d_par2 = nullptr, *d_dir = nullptr, *d_out = nullptr, *d_hv = nullptr;
    double *d_hess = nullptr, *d_Hbb = nullptr, *d_rhs = nullptr, *d_work = nullptr;
    int* d_info = nullptr;
    int lwork = 0;
    std::vector<double> par, out, grad, H_cols, L, step, gb;
    ssde_laplace_opts opts{};
    bool factor_valid = false;       // d_Hbb holds the Cholesky factor of H_bb at the mode of the previous evaluation
    std::string err;
    ~ssde_laplace() {
        cudaSetDevice(device);
        if (cs) cusolverDnDestroy(cs);
        for (double* p : {d_par, d_par2, d_dir, d_out, d_hv, d_hess, d_Hbb, d_rhs, d_work}) if (p) cudaFree(p);
        if (d_info) cudaFree(d_info);
    }
};

namespace {

#define LP_CUDA(expr)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) { w->err = std::string(#expr) + ": " + cudaGetErrorString(e__); return SSDE_ERR_CUDA; } \
    } while (0)
#define LP_SOLVER(expr)                                                                      \
    do {                                                                                     \
        cusolverStatus_t s__ = (expr);                                                       \
        if (s__ != CUSOLVER_STATUS_SUCCESS) { w->err = std::string(#expr) + " failed with cuSOLVER status " + std::to_string((int)s__); return SSDE_ERR_CUDA; } \
    } while (0)
#define LP_TRY(expr)                                                                         \
    do {                                                                                     \
        int rc__ = (expr);                                                                   \
        if (rc__ != SSDE_OK) { if (w->err.empty()) w->err = ssde_last_error(w->h); return rc__; } \
    } while (0)

void default_opts(ssde_laplace_opts& o) {
    o.max_newton = 100;
    o.grad_tol = 1e-8;
    o.fd_step = 1e-3;
    o.richardson = 1;
}

// joint value only at w->par
int joint_value(ssde_laplace* w, const std::vector<double>& par, double& v) {
    LP_CUDA(cudaMemcpyAsync(w->d_par2, par.data(), sizeof(double) * w->np, cudaMemcpyHostToDevice, w->st));
    LP_TRY(ssde_eval_device(w->h, w->d_par2, 0, w->d_out, w->st));
    double o[1];
    LP_CUDA(cudaMemcpyAsync(o, w->d_out, sizeof(double), cudaMemcpyDeviceToHost, w->st));
    LP_CUDA(cudaStreamSynchronize(w->st));
    v = o[0];
    return SSDE_OK;
}

// joint value and gradient at `par` (one plain evaluation, no tangent pass)
int joint_grad(ssde_laplace* w, const std::vector<double>& par, double& v, std::vector<double>& g) {
    const int np = w->np;
    LP_CUDA(cudaMemcpyAsync(w->d_par2, par.data(), sizeof(double) * np, cudaMemcpyHostToDevice, w->st));
    LP_TRY(ssde_eval_device(w->h, w->d_par2, 1, w->d_out, w->st));
    std::vector<double> o(np + 2);
    LP_CUDA(cudaMemcpyAsync(o.data(), w->d_out, sizeof(double) * (np + 2), cudaMemcpyDeviceToHost, w->st));
    LP_CUDA(cudaStreamSynchronize(w->st));
    if (o[np + 1] != 0.0) {
        w->err = ((unsigned)o[np + 1] & 2u) ? "device-side failure: innovation variance F <= 0 in the filter" : "device-side failure (scan look-back timed out)";
        return SSDE_ERR_NUMERIC;
    }
    v = o[0];
    g.assign(o.begin() + 1, o.begin() + 1 + np);
    return SSDE_OK;
}

// value, gradient and the coeff_re columns of the joint Hessian at w->par (n_b tangent passes)
int joint_hess_cols(ssde_laplace* w, double& v) {
    const int np = w->np, nb = w->nb;
    LP_CUDA(cudaMemcpyAsync(w->d_par, w->par.data(), sizeof(double) * np, cudaMemcpyHostToDevice, w->st));
    LP_TRY(ssde_hess_cols_device(w->h, w->d_par, w->o_re, nb, w->d_out, w->d_hess, w->st));
    LP_CUDA(cudaMemcpyAsync(w->out.data(), w->d_out, sizeof(double) * (np + 2), cudaMemcpyDeviceToHost, w->st));
    LP_CUDA(cudaStreamSynchronize(w->st));
    if (w->out[np + 1] != 0.0) {
        w->err = ((unsigned)w->out[np + 1] & 2u) ? "device-side failure: innovation variance F <= 0 in the filter" : "device-side failure (scan look-back timed out)";
        return SSDE_ERR_NUMERIC;
    }
    v = w->out[0];
    std::memcpy(w->grad.data(), w->out.data() + 1, sizeof(double) * np);
    return SSDE_OK;
}

// H_bb (+ ridge I) -> Cholesky factor in d_Hbb (lower); info > 0: not positive definite
int factor_bb(ssde_laplace* w, double ridge, int& info) {
    const int np = w->np, nb = w->nb;
    // rows o_re .. o_re + nb - 1 of the nb columns of d_hess (leading dimension np)
    LP_CUDA(cudaMemcpy2DAsync(w->d_Hbb, sizeof(double) * nb, w->d_hess + w->o_re, sizeof(double) * np,
                              sizeof(double) * nb, nb, cudaMemcpyDeviceToDevice, w->st));
    if (ridge != 0.0) {
        std::vector<double> diag(nb);
        LP_CUDA(cudaMemcpy2DAsync(diag.data(), sizeof(double), w->d_Hbb, sizeof(double) * (nb + 1), sizeof(double), nb,
                                  cudaMemcpyDeviceToHost, w->st));
        LP_CUDA(cudaStreamSynchronize(w->st));
        for (double& d : diag) d += ridge;
        LP_CUDA(cudaMemcpy2DAsync(w->d_Hbb, sizeof(double) * (nb + 1), diag.data(), sizeof(double), sizeof(double), nb,
                                  cudaMemcpyHostToDevice, w->st));
    }
    LP_SOLVER(cusolverDnDpotrf(w->cs, CUBLAS_FILL_MODE_LOWER, nb, w->d_Hbb, nb, w->d_work, w->lwork, w->d_info));
    LP_CUDA(cudaMemcpyAsync(&info, w->d_info, sizeof(int), cudaMemcpyDeviceToHost, w->st));
    LP_CUDA(cudaStreamSynchronize(w->st));
    return SSDE_OK;
}

// x <- H_bb^-1 x with the current factor
int solve_bb(ssde_laplace* w, std::vector<double>& x) {
    const int nb = w->nb;
    int info = 0;
    LP_CUDA(cudaMemcpyAsync(w->d_rhs, x.data(), sizeof(double) * nb, cudaMemcpyHostToDevice, w->st));
    LP_SOLVER(cusolverDnDpotrs(w->cs, CUBLAS_FILL_MODE_LOWER, nb, 1, w->d_Hbb, nb, w->d_rhs, nb, w->d_info));
    LP_CUDA(cudaMemcpyAsync(x.data(), w->d_rhs, sizeof(double) * nb, cudaMemcpyDeviceToHost, w->st));
    LP_CUDA(cudaMemcpyAsync(&info, w->d_info, sizeof(int), cudaMemcpyDeviceToHost, w->st));
    LP_CUDA(cudaStreamSynchronize(w->st));
    if (info != 0) { w->err = "potrs failed"; return SSDE_ERR_NUMERIC; }
    return SSDE_OK;
}

// H(par_at) * dir  (one tangent pass); hv on the host
int hvp_at(ssde_laplace* w, const std::vector<double>& par_at, const std::vector<double>& dir, std::vector<double>& hv) {
    const int np = w->np;
    LP_CUDA(cudaMemcpyAsync(w->d_par2, par_at.data(), sizeof(double) * np, cudaMemcpyHostToDevice, w->st));
    LP_CUDA(cudaMemcpyAsync(w->d_dir, dir.data(), sizeof(double) * np, cudaMemcpyHostToDevice, w->st));
    LP_TRY(ssde_hvp_device(w->h, w->d_par2, w->d_dir, w->d_out, w->d_hv, w->st));
    LP_CUDA(cudaMemcpyAsync(hv.data(), w->d_hv, sizeof(double) * np, cudaMemcpyDeviceToHost, w->st));
    LP_CUDA(cudaStreamSynchronize(w->st));
    return SSDE_OK;
}

}  // namespace

extern "C" {

static int laplace_create_impl(ssde_handle* h, const ssde_laplace_opts* opts, ssde_laplace** out) {
    if (!h || !out) return SSDE_ERR_BAD_ARG;
    *out = nullptr;
    ssde_laplace* w = new (std::nothrow) ssde_laplace();
    if (!w) return SSDE_ERR_BAD_ARG;
    std::unique_ptr<ssde_laplace> guard(w);        // frees the workspace on any error or exception
    auto fail = [&](int rc) { return rc; };
    w->h = h;
    w->device = ssde_device(h);
    w->st = (cudaStream_t)ssde_stream(h);
    int32_t off[4], siz[4];
    ssde_par_layout(h, off, siz);
    w->np = ssde_n_par(h);
    w->o_re = off[3];
    w->nb = siz[3];
    default_opts(w->opts);
    if (opts) {
        if (opts->max_newton > 0) w->opts.max_newton = opts->max_newton;
        if (opts->grad_tol > 0) w->opts.grad_tol = opts->grad_tol;
        if (opts->fd_step > 0) w->opts.fd_step = opts->fd_step;
        w->opts.richardson = opts->richardson;
    }
    const int np = w->np, nb = std::max(w->nb, 1);
    if (cudaSetDevice(w->device) != cudaSuccess) return fail(SSDE_ERR_CUDA);
    bool ok = true;
    for (double** p : {&w->d_par, &w->d_par2, &w->d_dir, &w->d_hv}) ok = ok && cudaMalloc(p, sizeof(double) * np) == cudaSuccess;
    ok = ok && cudaMalloc(&w->d_out, sizeof(double) * (np + 2)) == cudaSuccess;
    ok = ok && cudaMalloc(&w->d_hess, sizeof(double) * (size_t)np * nb) == cudaSuccess;
    ok = ok && cudaMalloc(&w->d_Hbb, sizeof(double) * (size_t)nb * nb) == cudaSuccess;
    ok = ok && cudaMalloc(&w->d_rhs, sizeof(double) * nb) == cudaSuccess;
    ok = ok && cudaMalloc(&w->d_info, sizeof(int)) == cudaSuccess;
    if (!ok) return fail(SSDE_ERR_CUDA);
    if (cusolverDnCreate(&w->cs) != CUSOLVER_STATUS_SUCCESS) return fail(SSDE_ERR_CUDA);
    if (cusolverDnSetStream(w->cs, w->st) != CUSOLVER_STATUS_SUCCESS) return fail(SSDE_ERR_CUDA);
    if (cusolverDnDpotrf_bufferSize(w->cs, CUBLAS_FILL_MODE_LOWER, nb, w->d_Hbb, nb, &w->lwork) != CUSOLVER_STATUS_SUCCESS) return fail(SSDE_ERR_CUDA);
    if (cudaMalloc(&w->d_work, sizeof(double) * std::max(w->lwork, 1)) != cudaSuccess) return fail(SSDE_ERR_CUDA);
    w->par.resize(np); w->out.resize(np + 2); w->grad.resize(np);
    w->H_cols.resize((size_t)np * nb); w->L.resize((size_t)nb * nb); w->step.resize(nb); w->gb.resize(nb);
    *out = guard.release();
    return SSDE_OK;
}

int ssde_laplace_create(ssde_handle* h, const ssde_laplace_opts* opts, ssde_laplace** out) {
    try {                                           // no C++ exception crosses the C ABI
        return laplace_create_impl(h, opts, out);
    } catch (...) {
        if (out) *out = nullptr;
        return SSDE_ERR_BAD_ARG;
    }
}

void ssde_laplace_destroy(ssde_laplace* w) { delete w; }

const char* ssde_laplace_error(const ssde_laplace* w) { return w ? w->err.c_str() : ""; }

int ssde_laplace_eval(ssde_laplace* w, double* par, int order, double* value, double* grad, ssde_laplace_info* info) {
    if (!w || !par || !value) return SSDE_ERR_BAD_ARG;
    if (order < 0 || order > 1 || (order == 1 && !grad)) { w->err = "order must be 0 or 1 (with a gradient buffer)"; return SSDE_ERR_BAD_ARG; }
    w->err.clear();
    const int np = w->np, nb = w->nb, o = w->o_re;
    LP_CUDA(cudaSetDevice(w->device));
    std::memcpy(w->par.data(), par, sizeof(double) * np);
    ssde_laplace_info li{};
    if (nb == 0) {                               // no random effect: the marginal IS the joint objective
        double v;
        LP_TRY(ssde_eval(w->h, par, order, &v, grad, nullptr));
        *value = v;
        li.joint = v;
        if (info) *info = li;
        return SSDE_OK;
    }
    // ---- warm start: chord iterations with the factor of the previous mode.  Between two outer
    //      (BFGS) steps theta moves little, so H_bb^-1 of the last mode is a good preconditioner: each
    //      iteration costs one plain evaluation instead of the n_b tangent passes of a Hessian; the
    //      exact H_bb is then built once, at the converged mode, where the log-determinant needs it.
    double v = 0.0;
    if (w->factor_valid) {
        std::vector<double> g, gt, trial;
        LP_TRY(joint_grad(w, w->par, v, g));
        ++li.n_value;
        for (int it = 0; it < 12; ++it) {
            double gmax = 0.0;
            for (int i = 0; i < nb; ++i) { w->gb[i] = g[o + i]; gmax = std::max(gmax, std::fabs(w->gb[i])); }
            if (!(gmax == gmax) || gmax <= std::max(w->opts.grad_tol, 1e-12 * std::fabs(v))) break;
            w->step = w->gb;
            LP_TRY(solve_bb(w, w->step));
            trial = w->par;
            for (int i = 0; i < nb; ++i) trial[o + i] = w->par[o + i] - w->step[i];
            double vt = 0.0;
            LP_TRY(joint_grad(w, trial, vt, gt));
            ++li.n_value;
            double gmax_t = 0.0;
            for (int i = 0; i < nb; ++i) gmax_t = std::max(gmax_t, std::fabs(gt[o + i]));
            // keep the step only if it is a clear improvement; otherwise hand over to Newton
            if (!(vt == vt) || !(gmax_t == gmax_t) || vt > v + 1e-14 * std::fabs(v) || gmax_t > 0.5 * gmax) break;
            w->par = trial; v = vt; g.swap(gt);
        }
    }
    w->factor_valid = false;
    // ---- inner problem: Newton on coeff_re with the exact H_bb, backtracking on the joint objective
    LP_TRY(joint_hess_cols(w, v));
    li.n_hess = 1;
    int pd_info = 0;
    bool converged = false;
    for (int it = 0; it <= w->opts.max_newton; ++it) {
        double gmax = 0.0;
        for (int i = 0; i < nb; ++i) { w->gb[i] = w->grad[o + i]; gmax = std::max(gmax, std::fabs(w->gb[i])); }
        li.grad_max = gmax;
        if (!(gmax == gmax)) { w->err = "non-finite gradient in the inner problem"; return SSDE_ERR_NUMERIC; }
        // absolute tolerance, or -- for objectives of 1e6 and more, whose gradient carries rounding noise
        // far above 1e-8 -- relative to the value: max|g_b| <= 1e-12 |g| puts b_hat within ~1e-12 of the
        // mode (H_bb scales with |g| too) and is reachable in fp64
        const double tol_eff = std::max(w->opts.grad_tol, 1e-12 * std::fabs(v));
        if (gmax <= tol_eff) { converged = true; break; }
        if (it == w->opts.max_newton) break;
        // Levenberg ridge until H_bb + ridge I is positive definite
        double ridge = 0.0;
        for (int tries = 0; tries < 40; ++tries) {
            LP_TRY(factor_bb(w, ridge, pd_info));
            if (pd_info == 0) break;
            ridge = (ridge == 0.0) ? 1e-6 * (1.0 + gmax) : ridge * 10.0;
        }
        if (pd_info != 0) { w->err = "H_bb could not be made positive definite"; return SSDE_ERR_NUMERIC; }
        w->step = w->gb;
        LP_TRY(solve_bb(w, w->step));
        double slope = 0.0;
        for (int i = 0; i < nb; ++i) slope += w->gb[i] * w->step[i];
        std::vector<double> trial = w->par;
        double t = 1.0, vn = 0.0;
        bool ok = false;
        // near the mode a failing search means "working precision reached": do not halve 40 times there
        const int max_ls = (gmax <= 1e3 * tol_eff) ? 6 : 40;
        for (int ls = 0; ls < max_ls; ++ls, t *= 0.5) {
            for (int i = 0; i < nb; ++i) trial[o + i] = w->par[o + i] - t * w->step[i];
            LP_TRY(joint_value(w, trial, vn));
            ++li.n_value;
            if (vn == vn && vn <= v - 1e-4 * t * slope + 1e-14 * std::fabs(v)) { ok = true; break; }
        }
        if (!ok) { converged = gmax <= 1e3 * tol_eff; break; }      // no descent possible at working precision
        w->par = trial;
        LP_TRY(joint_hess_cols(w, v));
        ++li.n_hess;
        ++li.n_newton;
    }
    li.converged = converged ? 1 : 0;
    // ---- at the mode: factor, log-determinant
    LP_TRY(factor_bb(w, 0.0, pd_info));
    if (pd_info != 0) {
        *value = INFINITY;                       // not a minimum in b: the Laplace value is undefined
        li.joint = v;
        std::memcpy(par, w->par.data(), sizeof(double) * np);
        if (info) *info = li;
        w->err = "H_bb is not positive definite at the inner optimum";
        return SSDE_ERR_NUMERIC;
    }
    w->factor_valid = true;                      // d_Hbb = chol(H_bb) at the mode: preconditioner of the next call
    LP_CUDA(cudaMemcpyAsync(w->L.data(), w->d_Hbb, sizeof(double) * (size_t)nb * nb, cudaMemcpyDeviceToHost, w->st));
    LP_CUDA(cudaMemcpyAsync(w->H_cols.data(), w->d_hess, sizeof(double) * (size_t)np * nb, cudaMemcpyDeviceToHost, w->st));
    LP_CUDA(cudaStreamSynchronize(w->st));
    double logdet = 0.0;
    for (int i = 0; i < nb; ++i) logdet += 2.0 * std::log(w->L[(size_t)i * nb + i]);
    li.joint = v;
    li.logdet = logdet;
    *value = v + 0.5 * logdet - 0.5 * nb * std::log(2.0 * M_PI);
    std::memcpy(par, w->par.data(), sizeof(double) * np);
    if (order == 0) { if (info) *info = li; return SSDE_OK; }

    // ---- gradient: w_k = sum_j D^3 g[z_j, z_j, e_k],  Z = L^-T (column j solves L' z_j = e_j)
    std::vector<double> wv(np, 0.0), dir(np), pp(np), hp(np), hm(np), hp2(np), hm2(np), z(nb);
    const double eps = w->opts.fd_step;
    for (int j = 0; j < nb; ++j) {
        // back-substitution L' z = e_j  (L lower, column-major: L[i + nb*c])
        for (int i = nb - 1; i >= 0; --i) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int c = i + 1; c < nb; ++c) s -= w->L[(size_t)i * nb + c] * z[c];      // L'(i,c) = L(c,i)
            z[i] = s / w->L[(size_t)i * nb + i];
        }
        double nrm = 0.0;
        for (int i = 0; i < nb; ++i) nrm = std::max(nrm, std::fabs(z[i]));
        if (nrm == 0.0) continue;
        std::fill(dir.begin(), dir.end(), 0.0);
        for (int i = 0; i < nb; ++i) dir[o + i] = z[i] / nrm;
        auto shifted = [&](double t, std::vector<double>& hv) {
            pp = w->par;
            for (int i = 0; i < nb; ++i) pp[o + i] += t * dir[o + i];
            return hvp_at(w, pp, dir, hv);
        };
        LP_TRY(shifted(eps, hp));
        LP_TRY(shifted(-eps, hm));
        li.n_hvp += 2;
        if (w->opts.richardson) {
            LP_TRY(shifted(0.5 * eps, hp2));
            LP_TRY(shifted(-0.5 * eps, hm2));
            li.n_hvp += 2;
        }
        const double sc = nrm * nrm;
        for (int k = 0; k < np; ++k) {
            double d = (hp[k] - hm[k]) / (2.0 * eps);
            if (w->opts.richardson) d = (4.0 * (hp2[k] - hm2[k]) / eps - d) / 3.0;
            wv[k] += sc * d;
        }
    }
    std::vector<double> u(nb);
    for (int i = 0; i < nb; ++i) u[i] = wv[o + i];
    LP_TRY(solve_bb(w, u));
    for (int k = 0; k < np; ++k) {
        if (k >= o && k < o + nb) { grad[k] = 0.0; continue; }     // random effects are integrated out
        double corr = 0.0;
        for (int j = 0; j < nb; ++j) corr += w->H_cols[(size_t)j * np + k] * u[j];
        grad[k] = w->grad[k] + 0.5 * (wv[k] - corr);
    }
    if (info) *info = li;
    return SSDE_OK;
}

int ssde_laplace_hessian_bb(ssde_laplace* w, double* hess_bb) {
    if (!w || !hess_bb) return SSDE_ERR_BAD_ARG;
    const int np = w->np, nb = w->nb, o = w->o_re;
    for (int j = 0; j < nb; ++j)
        for (int i = 0; i < nb; ++i) hess_bb[(size_t)j * nb + i] = w->H_cols[(size_t)j * np + o + i];
    return SSDE_OK;
}

}  // extern "C"
