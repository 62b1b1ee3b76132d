/* TEST INFRASTRUCTURE ONLY.  Declarations of the part of R's C API that r_shim/ uses, so that the
 * shim can be type-checked (g++ -fsyntax-only) in an image without R.  Signatures follow R's
 * public Rinternals.h; nothing here is linked or shipped. */
#ifndef R_STUB_RINTERNALS_H
#define R_STUB_RINTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
/* R_ext/Arith.h */
extern double R_PosInf, R_NaReal;
#define NA_REAL R_NaReal
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
double* REAL(SEXP);
int* INTEGER(SEXP);
const char* CHAR(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
R_xlen_t Rf_xlength(SEXP);
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_coerceVector(SEXP, SEXPTYPE);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_install(const char*);
SEXP Rf_mkChar(const char*);
SEXP Rf_duplicate(SEXP);
SEXP Rf_ScalarInteger(int);
int Rf_asInteger(SEXP);
SEXP R_do_slot(SEXP, SEXP);
void Rf_error(const char*, ...) __attribute__((noreturn));
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#ifdef __cplusplus
}
#endif
#endif
