// R <-> libsmoothsde_b200 shim.  NOT compiled in this image (no R headers here); it is the
// binding a smoothSDE maintainer adds next to src/smoothSDE.cpp.  It replaces the TMB entry
// points that R/sde.R reaches through TMB::MakeADFun (registered in src/init.c:6-15):
//   MakeADFunObject(data, parameters, reportenv, control)  ->  ssde_make(data, device)
//   EvalADFunObject(ptr, theta, control)                   ->  ssde_fn_gr(ptr, par, order)
//   REPORT(aest_all) via the report environment            ->  ssde_aest(ptr)
//   MakeADHessObject2 (obj$he, the Laplace inner problem)  ->  ssde_he(ptr, par), ssde_laplace_new /
//                                                              ssde_laplace_fn_gr(lap, par, order)
// `map` / `random` stay on the R side (r_shim/R/adfun.R), exactly as in TMB's own R layer.
#include <R.h>
#include <Rinternals.h>
#include <string.h>

#include "smoothsde_b200.h"

namespace {

SEXP list_get(SEXP lst, const char* name) {
    SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
    for (R_xlen_t i = 0; i < Rf_xlength(lst); ++i)
        if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(lst, i);
    return R_NilValue;
}

// dgTMatrix (as_sparse(), R/utility.R:204-213): slots i, j (0-based int), x, Dim
ssde_triplet triplet_of(SEXP m) {
    ssde_triplet t;
    SEXP dim = R_do_slot(m, Rf_install("Dim"));
    t.nrow = INTEGER(dim)[0];
    t.ncol = INTEGER(dim)[1];
    SEXP x = R_do_slot(m, Rf_install("x"));
    t.nnz = Rf_xlength(x);
    t.i = INTEGER(R_do_slot(m, Rf_install("i")));
    t.j = INTEGER(R_do_slot(m, Rf_install("j")));
    t.x = REAL(x);
    return t;
}

void finalizer(SEXP ptr) {
    ssde_handle* h = (ssde_handle*)R_ExternalPtrAddr(ptr);
    if (h) { ssde_destroy(h); R_ClearExternalPtr(ptr); }
}

int model_of(const char* type) {
    if (!strcmp(type, "BM")) return SSDE_BM;
    if (!strcmp(type, "OU")) return SSDE_OU;
    if (!strcmp(type, "CTCRW")) return SSDE_CTCRW;
    if (!strcmp(type, "BM_SSM")) return SSDE_BM_SSM;
    if (!strcmp(type, "OU_SSM")) return SSDE_OU_SSM;
    return -1;          // ssde_create answers SSDE_ERR_UNKNOWN_TYPE: "Unknown SDE type" (smoothSDE.cpp:25)
}

}  // namespace

extern "C" {

// data = the tmb_dat list of SDE$setup() (R/sde.R:528-598)
SEXP ssde_make(SEXP data, SEXP device) {
    ssde_desc d;
    memset(&d, 0, sizeof(d));
    d.model = model_of(CHAR(STRING_ELT(list_get(data, "type"), 0)));
    SEXP obs = list_get(data, "obs");
    SEXP dim = Rf_getAttrib(obs, R_DimSymbol);
    d.n = INTEGER(dim)[0];
    d.n_dim = INTEGER(dim)[1];
    // data$ID is a factor: integer codes -> doubles (TMB's DATA_VECTOR does the same)
    SEXP id = PROTECT(Rf_coerceVector(list_get(data, "ID"), REALSXP));
    d.ID = REAL(id);
    d.times = REAL(list_get(data, "times"));
    d.obs = REAL(obs);                                  // column-major, NA_real_ is a NaN
    d.X_fe = triplet_of(list_get(data, "X_fe"));
    d.X_re = triplet_of(list_get(data, "X_re"));
    d.S = triplet_of(list_get(data, "S"));
    SEXP ncol_re = PROTECT(Rf_coerceVector(list_get(data, "ncol_re"), INTSXP));
    d.n_smooth = (int32_t)Rf_xlength(ncol_re);
    d.ncol_re = INTEGER(ncol_re);
    d.include_penalty = Rf_asInteger(list_get(data, "include_penalty"));
    SEXP a0 = list_get(data, "a0");
    if (a0 != R_NilValue) {
        d.a0 = REAL(a0);
        d.n_ID = INTEGER(Rf_getAttrib(a0, R_DimSymbol))[0];
        d.P0 = REAL(list_get(data, "P0"));
        SEXP H = list_get(data, "H_array");
        if (H != R_NilValue && Rf_xlength(H) > 1) { d.H_array = REAL(H); d.H_len = Rf_xlength(H); }
    }
    // decay terms (nllk_sde.hpp:31-33): R passes t_decay = col_decay = ind_decay = 0 when there are none (R/sde.R:645-647)
    SEXP td = list_get(data, "t_decay");
    SEXP cd = PROTECT(Rf_coerceVector(td != R_NilValue ? list_get(data, "col_decay") : Rf_ScalarInteger(0), INTSXP));
    SEXP id_ = PROTECT(Rf_coerceVector(td != R_NilValue ? list_get(data, "ind_decay") : Rf_ScalarInteger(0), INTSXP));
    if (td != R_NilValue && Rf_xlength(td) > 1) {
        d.t_decay = REAL(td); d.t_decay_len = Rf_xlength(td);
        d.col_decay = INTEGER(cd); d.ind_decay = INTEGER(id_); d.n_col_decay = (int32_t)Rf_xlength(cd);
    }
    d.device = Rf_asInteger(device);
    ssde_handle* h = NULL;
    int rc = ssde_create(&d, &h);
    UNPROTECT(4);
    if (rc != SSDE_OK) Rf_error("smoothsde_b200: %s", ssde_create_error());
    SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

// list(value = nllk, gradient = d nllk / d par) for the FULL parameter vector
SEXP ssde_fn_gr(SEXP ptr, SEXP par, SEXP order) {
    ssde_handle* h = (ssde_handle*)R_ExternalPtrAddr(ptr);
    if (!h) Rf_error("smoothsde_b200: handle was freed");
    const int np = ssde_n_par(h), ord = Rf_asInteger(order);
    if (Rf_xlength(par) != np) Rf_error("smoothsde_b200: parameter vector has length %d, expected %d", (int)Rf_xlength(par), np);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP val = PROTECT(Rf_allocVector(REALSXP, 1));
    SEXP grad = PROTECT(Rf_allocVector(REALSXP, ord >= 1 ? np : 0));
    int rc = ssde_eval(h, REAL(par), ord, REAL(val), ord >= 1 ? REAL(grad) : NULL, NULL);
    if (rc != SSDE_OK) { UNPROTECT(3); Rf_error("smoothsde_b200: %s", ssde_last_error(h)); }
    SET_VECTOR_ELT(out, 0, val);
    SET_VECTOR_ELT(out, 1, grad);
    SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_STRING_ELT(nm, 0, Rf_mkChar("value"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("gradient"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(4);
    return out;
}

// offsets[5] / sizes[5] of log_sigma_obs, coeff_fe, log_lambda, log_decay, coeff_re in the full
// vector (the PARAMETER order of the templates, nllk_ctcrw.hpp:135-140 / nllk_sde.hpp:42-45)
SEXP ssde_layout(SEXP ptr) {
    ssde_handle* h = (ssde_handle*)R_ExternalPtrAddr(ptr);
    SEXP out = PROTECT(Rf_allocVector(INTSXP, 10));
    int32_t off[4], siz[4], od = -1, nd = 0;
    ssde_par_layout(h, off, siz);
    ssde_decay_layout(h, &od, &nd);
    int* o = INTEGER(out);
    o[0] = off[0]; o[1] = off[1]; o[2] = off[2]; o[3] = od; o[4] = off[3];
    o[5] = siz[0]; o[6] = siz[1]; o[7] = siz[2]; o[8] = nd; o[9] = siz[3];
    UNPROTECT(1);
    return out;
}

// REPORT(aest_all) (nllk_ctcrw.hpp:249) at the parameters of the last evaluation
// n_state = columns of aest_all: 2 * n_dim for CTCRW, n_dim for BM_SSM / OU_SSM (= ncol(a0))
SEXP ssde_aest(SEXP ptr, SEXP n, SEXP n_state) {
    ssde_handle* h = (ssde_handle*)R_ExternalPtrAddr(ptr);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, Rf_asInteger(n), Rf_asInteger(n_state)));
    int rc = ssde_report(h, REAL(out));
    UNPROTECT(1);
    if (rc != SSDE_OK) Rf_error("smoothsde_b200: %s", ssde_last_error(h));
    return out;
}

// obj$he(par): joint Hessian [n_par x n_par] of the penalised objective (R/sde.R:1363)
SEXP ssde_he(SEXP ptr, SEXP par) {
    ssde_handle* h = (ssde_handle*)R_ExternalPtrAddr(ptr);
    if (!h) Rf_error("smoothsde_b200: handle was freed");
    const int np = ssde_n_par(h);
    if (Rf_xlength(par) != np) Rf_error("smoothsde_b200: parameter vector has length %d, expected %d", (int)Rf_xlength(par), np);
    SEXP H = PROTECT(Rf_allocMatrix(REALSXP, np, np));
    SEXP g = PROTECT(Rf_allocVector(REALSXP, np));
    double v;
    int rc = ssde_eval(h, REAL(par), 2, &v, REAL(g), REAL(H));
    UNPROTECT(2);
    if (rc != SSDE_OK) Rf_error("smoothsde_b200: %s", ssde_last_error(h));
    return H;
}

// Laplace-marginal object (random = "coeff_re", R/sde.R:522-524): workspace tied to a handle
static void lap_finalizer(SEXP ptr) {
    ssde_laplace* w = (ssde_laplace*)R_ExternalPtrAddr(ptr);
    if (w) { ssde_laplace_destroy(w); R_ClearExternalPtr(ptr); }
}
SEXP ssde_laplace_new(SEXP ptr) {
    ssde_handle* h = (ssde_handle*)R_ExternalPtrAddr(ptr);
    ssde_laplace* w = NULL;
    if (ssde_laplace_create(h, NULL, &w) != SSDE_OK) Rf_error("smoothsde_b200: ssde_laplace_create failed");
    SEXP out = PROTECT(R_MakeExternalPtr(w, R_NilValue, ptr));      // keeps the engine handle alive
    R_RegisterCFinalizerEx(out, lap_finalizer, TRUE);
    UNPROTECT(1);
    return out;
}
// list(value = f(theta), gradient [full length, 0 in the coeff_re slots], par = full vector with
// coeff_re at the inner optimum); `par` carries the starting value of coeff_re
SEXP ssde_laplace_fn_gr(SEXP lap, SEXP par, SEXP order) {
    ssde_laplace* w = (ssde_laplace*)R_ExternalPtrAddr(lap);
    if (!w) Rf_error("smoothsde_b200: Laplace workspace was freed");
    const int np = (int)Rf_xlength(par), ord = Rf_asInteger(order);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SEXP val = PROTECT(Rf_allocVector(REALSXP, 1));
    SEXP grad = PROTECT(Rf_allocVector(REALSXP, ord >= 1 ? np : 0));
    SEXP p = PROTECT(Rf_duplicate(par));
    // outputs are defined on EVERY path: ssde_laplace_eval writes *value only when it gets that far
    // and never writes the gradient on an error path
    REAL(val)[0] = R_PosInf;
    for (int i = 0; i < (ord >= 1 ? np : 0); ++i) REAL(grad)[i] = NA_REAL;
    int rc = ssde_laplace_eval(w, REAL(p), ord, REAL(val), ord >= 1 ? REAL(grad) : NULL, NULL);
    // the only failure an optimiser may step away from: H_bb not positive definite at the inner
    // optimum (value = Inf, as TMB's inner problem reports); anything else is an error
    const bool not_pd = rc == SSDE_ERR_NUMERIC && strstr(ssde_laplace_error(w), "not positive definite at the inner optimum") != NULL;
    if (rc != SSDE_OK && !not_pd) { UNPROTECT(4); Rf_error("smoothsde_b200: %s", ssde_laplace_error(w)); }
    if (rc != SSDE_OK) {                           // not_pd: Inf value, NA gradient, starting coeff_re kept
        REAL(val)[0] = R_PosInf;
        for (int i = 0; i < (ord >= 1 ? np : 0); ++i) REAL(grad)[i] = NA_REAL;
        memcpy(REAL(p), REAL(par), sizeof(double) * (size_t)np);
    }
    SET_VECTOR_ELT(out, 0, val);
    SET_VECTOR_ELT(out, 1, grad);
    SET_VECTOR_ELT(out, 2, p);
    SEXP nm = PROTECT(Rf_allocVector(STRSXP, 3));
    SET_STRING_ELT(nm, 0, Rf_mkChar("value"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("gradient"));
    SET_STRING_ELT(nm, 2, Rf_mkChar("par"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(5);
    return out;
}

SEXP ssde_free(SEXP ptr) { finalizer(ptr); return R_NilValue; }

}  // extern "C"
