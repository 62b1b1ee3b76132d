/* TEST INFRASTRUCTURE ONLY: see ../Rinternals.h. */
#ifndef R_STUB_RDYNLOAD_H
#define R_STUB_RDYNLOAD_H
typedef void* (*DL_FUNC)();
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
int R_registerRoutines(DllInfo*, const void*, const R_CallMethodDef*, const void*, const void*);
int R_useDynamicSymbols(DllInfo*, int);
#endif
