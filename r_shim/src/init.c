/* Registration table of the shim -- takes the place of the TMB table in src/init.c:6-26 of the
 * reference (MakeADFunObject, EvalADFunObject, ... -> the routines of shim.cpp).  Loaded through
 * useDynLib(smoothSDE, .registration = TRUE) (NAMESPACE:33) as before. */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

extern SEXP ssde_make(SEXP, SEXP);
extern SEXP ssde_fn_gr(SEXP, SEXP, SEXP);
extern SEXP ssde_layout(SEXP);
extern SEXP ssde_aest(SEXP, SEXP, SEXP);
extern SEXP ssde_he(SEXP, SEXP);
extern SEXP ssde_laplace_new(SEXP);
extern SEXP ssde_laplace_fn_gr(SEXP, SEXP, SEXP);
extern SEXP ssde_free(SEXP);

static const R_CallMethodDef CallEntries[] = {
    {"ssde_make",   (DL_FUNC) &ssde_make,   2},
    {"ssde_fn_gr",  (DL_FUNC) &ssde_fn_gr,  3},
    {"ssde_layout", (DL_FUNC) &ssde_layout, 1},
    {"ssde_aest",   (DL_FUNC) &ssde_aest,   3},
    {"ssde_he",     (DL_FUNC) &ssde_he,     2},
    {"ssde_laplace_new",   (DL_FUNC) &ssde_laplace_new,   1},
    {"ssde_laplace_fn_gr", (DL_FUNC) &ssde_laplace_fn_gr, 3},
    {"ssde_free",   (DL_FUNC) &ssde_free,   1},
    {NULL, NULL, 0}
};

void R_init_smoothSDE(DllInfo *dll) {
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
