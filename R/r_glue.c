/* R/r_glue.c -- optional .Call glue over libnngp_b200.so: external-pointer contexts with a finalizer, and
 * entry points whose outputs may be long vectors (more than 2^31 - 1 elements), which .C() cannot carry.
 *
 * Compile where R is installed (needs only R's own headers; link against the library):
 *     R CMD SHLIB R/r_glue.c -Iinclude -Llib -lnngp_b200 -o lib/nngp_b200_r.so
 * and load with dyn.load("lib/libnngp_b200.so"); dyn.load("lib/nngp_b200_r.so").  The build image has no R: there this file is
 * only syntax-checked against a stub of the handful of R API declarations it uses (tests/test_r_glue_cpu.py).
 *
 * Rules followed (SURVEY.md 8b "Ownership"): every SEXP that is allocated is PROTECTed until returned; no raw pointer into an R
 * object is kept past the call; Rf_error is only raised from this C frame after the library call has returned (the library never
 * longjmps). */
#include <R.h>
#include <Rinternals.h>

#include "nngp_b200.h"

static void check_status(int status) {
    if (status != NNGP_OK) {
        char msg[1024];
        int len = (int)sizeof(msg);
        nngp_last_error(msg, &len);
        Rf_error("libnngp_b200 status %d: %s", status, msg);
    }
}

static void ctx_finalizer(SEXP ptr) {
    int *id = (int *)R_ExternalPtrAddr(ptr);
    if (id) {
        int status = 0;
        nngp_ctx_destroy(id, &status);   /* errors are ignored in a finalizer */
        R_Free(id);
        R_ClearExternalPtr(ptr);
    }
}

static int ctx_of(SEXP ptr) {
    int *id = (int *)R_ExternalPtrAddr(ptr);
    if (!id) Rf_error("libnngp_b200: this context has been destroyed");
    return *id;
}

/* .Call("nngp_r_ctx_create", locs, NNarray, coloring, locs_match, covfun_id, device, layout) -> external pointer; the context is
 * destroyed when the pointer is garbage collected (or by nngp_r_ctx_destroy) */
SEXP nngp_r_ctx_create(SEXP locs, SEXP NNarray, SEXP coloring, SEXP locs_match, SEXP covfun_id, SEXP device, SEXP layout) {
    if (!Rf_isReal(locs) || !Rf_isMatrix(locs) || !Rf_isInteger(NNarray) || !Rf_isMatrix(NNarray) || !Rf_isInteger(coloring) || !Rf_isInteger(locs_match))
        Rf_error("nngp_r_ctx_create: locs must be a double matrix, NNarray an integer matrix, coloring / locs_match integer vectors");
    int n = Rf_nrows(locs), d = Rf_ncols(locs), m = Rf_ncols(NNarray) - 1, n_obs = (int)XLENGTH(locs_match);
    int cov = Rf_asInteger(covfun_id), dev = Rf_asInteger(device), lay = Rf_asInteger(layout), id = -1, status = 0;
    if (Rf_nrows(NNarray) != n || XLENGTH(coloring) != n) Rf_error("nngp_r_ctx_create: NNarray / coloring do not match locs");
    nngp_ctx_create(&n, &d, &m, REAL(locs), INTEGER(NNarray), INTEGER(coloring), &n_obs, INTEGER(locs_match), &cov, &dev, &lay, &id, &status);
    check_status(status);
    int *slot = R_Calloc(1, int);
    *slot = id;
    SEXP ptr = PROTECT(R_MakeExternalPtr(slot, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

SEXP nngp_r_ctx_destroy(SEXP ptr) {
    ctx_finalizer(ptr);
    return R_NilValue;
}

SEXP nngp_r_ctx_id(SEXP ptr) { return Rf_ScalarInteger(ctx_of(ptr)); }

/* .Call("nngp_r_chains_run", list_of_ctx, params (n_chains x (5 + k), rows = chains), n_iter, thin, n_chromatic, iter_start,
 *        chain_index, rng_mode, var_y, max_concurrent, keep_field)
 * -> list(params = matrix, records = list of n_iter x (3 + k) matrices, field = list of n_frec x n_locs matrices or NULL).
 * The field records of all chains are written by the library into ONE buffer, which may be a long vector. */
SEXP nngp_r_chains_run(SEXP ctxs, SEXP params, SEXP n_iter_, SEXP thin_, SEXP n_chromatic_, SEXP iter_start_, SEXP chain_index, SEXP rng_mode_,
                       SEXP var_y_, SEXP max_concurrent_, SEXP keep_field_, SEXP n_locs_) {
    if (!Rf_isNewList(ctxs) || !Rf_isReal(params) || !Rf_isMatrix(params) || !Rf_isInteger(chain_index))
        Rf_error("nngp_r_chains_run: ctxs must be a list of contexts, params a double matrix, chain_index an integer vector");
    int nc = (int)XLENGTH(ctxs), k = Rf_ncols(params) - 5, n_iter = Rf_asInteger(n_iter_), n_chromatic = Rf_asInteger(n_chromatic_);
    int iter_start = Rf_asInteger(iter_start_), rng_mode = Rf_asInteger(rng_mode_), max_conc = Rf_asInteger(max_concurrent_);
    int keep = Rf_asLogical(keep_field_), n_locs = Rf_asInteger(n_locs_), status = 0;
    double thin = Rf_asReal(thin_), var_y = Rf_asReal(var_y_);
    if (Rf_nrows(params) != nc || XLENGTH(chain_index) != nc || k < 1 || n_iter < 0) Rf_error("nngp_r_chains_run: inconsistent sizes");
    R_xlen_t n_frec = (R_xlen_t)nearbyint(n_iter * thin);
    int *ids = (int *)R_alloc(nc, sizeof(int));
    for (int c = 0; c < nc; c++) ids[c] = ctx_of(VECTOR_ELT(ctxs, c));
    /* chain-major copies of the per-chain parameter rows (R matrices are column-major) */
    double *p = (double *)R_alloc((size_t)nc * (5 + k), sizeof(double));
    for (int c = 0; c < nc; c++)
        for (int j = 0; j < 5 + k; j++) p[(size_t)c * (5 + k) + j] = REAL(params)[c + (size_t)nc * j];
    SEXP rec = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)nc * n_iter * (3 + k)));
    SEXP frec = PROTECT(keep ? Rf_allocVector(REALSXP, (R_xlen_t)nc * (n_frec > 0 ? n_frec : 1) * n_locs) : R_NilValue);
    SEXP acc = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)nc * 2 * n_iter));
    nngp_chains_run(&nc, ids, &k, p, &n_iter, &thin, &n_chromatic, &iter_start, INTEGER(chain_index), &rng_mode, &var_y, &max_conc, REAL(rec),
                    keep ? REAL(frec) : NULL, INTEGER(acc), &status);
    if (status != NNGP_OK) { UNPROTECT(3); check_status(status); }
    SEXP pout = PROTECT(Rf_allocMatrix(REALSXP, nc, 5 + k));
    for (int c = 0; c < nc; c++)
        for (int j = 0; j < 5 + k; j++) REAL(pout)[c + (size_t)nc * j] = p[(size_t)c * (5 + k) + j];
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
    SET_VECTOR_ELT(out, 0, pout);
    SET_VECTOR_ELT(out, 1, rec);     /* chain after chain: n_iter x (3 + k) blocks, column-major */
    SET_VECTOR_ELT(out, 2, frec);    /* chain after chain: n_frec x n_locs blocks, column-major (R's records$field layout) */
    SET_VECTOR_ELT(out, 3, acc);
    UNPROTECT(5);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"nngp_r_ctx_create", (DL_FUNC)&nngp_r_ctx_create, 7},
    {"nngp_r_ctx_destroy", (DL_FUNC)&nngp_r_ctx_destroy, 1},
    {"nngp_r_ctx_id", (DL_FUNC)&nngp_r_ctx_id, 1},
    {"nngp_r_chains_run", (DL_FUNC)&nngp_r_chains_run, 12},
    {NULL, NULL, 0}};

void R_init_nngp_b200_r(DllInfo *dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, TRUE);
}
