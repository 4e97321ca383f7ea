/* Minimal stand-in for R's C API, ONLY so that `gcc -fsyntax-only` can type-check r-shim/src/varycoef_b200_r.c in an
 * image without R (tests/test_host_logic.py).  Not R, not shipped, declares just what the shim uses. */
#ifndef VCB_CHECK_R_H
#define VCB_CHECK_R_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
extern SEXP R_NilValue;
extern double R_NaReal;
#define NA_REAL R_NaReal
void Rf_error(const char*, ...);
double* REAL(SEXP);
int* INTEGER(SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int Rf_asLogical(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_length(SEXP);
SEXP Rf_coerceVector(SEXP, unsigned int);
int Rf_isNull(SEXP);
SEXP Rf_allocVector(unsigned int, ptrdiff_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_ScalarReal(double);
SEXP Rf_install(const char*);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP SET_VECTOR_ELT(SEXP, ptrdiff_t, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
int R_registerRoutines(DllInfo*, const void*, const R_CallMethodDef*, const void*, const void*);
Rboolean R_useDynamicSymbols(DllInfo*, Rboolean);
#endif
