/* Stub of the R API declarations R/r_glue.c uses, for a syntax / prototype check in an image without R (tests/test_r_glue_cpu.py).
 * Signatures follow R's public headers (Rinternals.h, R_ext/Rdynload.h, R_ext/RS.h, R_ext/Memory.h); nothing here is linked. */
#ifndef NNGP_R_STUB_H
#define NNGP_R_STUB_H
#include <math.h>
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
#define REALSXP 14
#define INTSXP 13
#define VECSXP 19
extern SEXP R_NilValue;
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
typedef void (*R_CFinalizer_t)(SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char *, ...) __attribute__((noreturn));
Rboolean Rf_isReal(SEXP), Rf_isInteger(SEXP), Rf_isMatrix(SEXP), Rf_isNewList(SEXP);
int Rf_nrows(SEXP), Rf_ncols(SEXP), Rf_asInteger(SEXP), Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
R_xlen_t XLENGTH(SEXP);
double *REAL(SEXP);
int *INTEGER(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t), SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t), Rf_allocMatrix(unsigned int, int, int), Rf_ScalarInteger(int);
SEXP R_MakeExternalPtr(void *, SEXP, SEXP);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
char *R_alloc(size_t, int);
void *R_chk_calloc(size_t, size_t);
void R_chk_free(void *);
#define R_Calloc(n, t) ((t *)R_chk_calloc((size_t)(n), sizeof(t)))
#define R_Free(p) (R_chk_free((void *)(p)), (p) = NULL)
int R_registerRoutines(DllInfo *, const void *, const R_CallMethodDef *, const void *, const void *);
Rboolean R_useDynamicSymbols(DllInfo *, Rboolean);
#endif
