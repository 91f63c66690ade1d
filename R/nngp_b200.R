# R/nngp_b200.R -- thin .C() glue over libnngp_b200.so (C ABI: include/nngp_b200.h).
#
# R is not available in the build image, so this file is untested there; it is the binding a maintainer of the reference
# would add (INTEGRATION.md).  Every ABI function takes only pointers and returns void, so .C() needs no compiled shim.
# NAOK = TRUE lets NA_integer_ (INT_MIN) in NNarray reach the library unchanged.

# The library is built to <repo>/lib/libnngp_b200.so (python <package>/build.py).  NNGP_B200_HOME, if set, is the repo root;
# otherwise the path is resolved relative to this file when it was source()d with chdir / from the repo root.
nngp_b200_home = function()
{
  home = Sys.getenv("NNGP_B200_HOME", unset = NA)
  if(!is.na(home)) return(home)
  this = tryCatch(normalizePath(sys.frame(1)$ofile), error = function(e) NA)   # set while source()ing this file
  if(!is.na(this)) return(dirname(dirname(this)))
  getwd()
}
.nngp_b200_home = nngp_b200_home()

nngp_b200_load = function(path = file.path(.nngp_b200_home, "lib", "libnngp_b200.so"))
{
  if(!is.loaded("nngp_ctx_create")) dyn.load(path)
  invisible(TRUE)
}

nngp_covfun_id = function(stationary_covfun)
{
  ids = c(exponential_isotropic = 0L, exponential_sphere = 1L, exponential_scaledim = 2L, exponential_spacetime = 3L,
          matern_isotropic = 4L, matern_sphere = 5L, matern_scaledim = 6L, matern_spacetime = 7L)
  if(!stationary_covfun %in% names(ids)) stop(paste("unknown stationary_covfun", stationary_covfun))
  ids[[stationary_covfun]]
}

nngp_check = function(res)
{
  if(res$status != 0L)
  {
    # .C() passes a character vector as char **: nngp_last_error_r writes into the first string (1024 blanks = room for the message)
    msg = .C("nngp_last_error_r", buf = strrep(" ", 1024), len = 1024L)$buf
    stop(paste0("libnngp_b200 status ", res$status, ": ", trimws(msg)))
  }
  res
}

# one context per chain: graph structure of vecchia_approx (Scripts/mcmc_nngp_initialize.R:80-110) uploaded once
nngp_ctx_create = function(locs, vecchia_approx, stationary_covfun, device = 0L, layout = 2L)
{
  locs = as.matrix(locs)
  res = nngp_check(.C("nngp_ctx_create", n = nrow(locs), d = ncol(locs), m = as.integer(ncol(vecchia_approx$NNarray) - 1L),
                      locs = as.double(locs), NNarray = as.integer(vecchia_approx$NNarray),
                      coloring = as.integer(vecchia_approx$coloring), n_obs = as.integer(vecchia_approx$n_obs),
                      locs_match = as.integer(vecchia_approx$locs_match), covfun_id = nngp_covfun_id(stationary_covfun),
                      device = as.integer(device), layout = as.integer(layout), ctx_id = integer(1), status = integer(1), NAOK = TRUE))
  res$ctx_id
}

nngp_ctx_destroy = function(ctx) invisible(.C("nngp_ctx_destroy", ctx_id = as.integer(ctx), status = integer(1)))

# Contexts are kept between the cycles of mcmc_nngp_run (one per chain: structure upload, tile build and colouring check at
# n = 1M cost more than a whole cycle's compute): an environment keyed by chain, reused while the structure is the same object
.nngp_b200_cache = new.env()
nngp_chain_context = function(chain, locs, vecchia_approx, stationary_covfun, device)
{
  key = paste0("chain_", chain)
  sig = list(n = nrow(as.matrix(locs)), m = ncol(vecchia_approx$NNarray) - 1L, covfun = stationary_covfun, device = as.integer(device),
             nn_head = vecchia_approx$NNarray[seq_len(min(64L, length(vecchia_approx$NNarray)))], n_obs = vecchia_approx$n_obs)
  hit = .nngp_b200_cache[[key]]
  if(!is.null(hit) && identical(hit$sig, sig)) return(hit$ctx)
  if(!is.null(hit)) nngp_ctx_destroy(hit$ctx)
  ctx = nngp_ctx_create(locs, vecchia_approx, stationary_covfun, device = device)
  .nngp_b200_cache[[key]] = list(sig = sig, ctx = ctx, regressors = FALSE)
  ctx
}
nngp_b200_release = function()
{
  for(key in ls(.nngp_b200_cache)) { nngp_ctx_destroy(.nngp_b200_cache[[key]]$ctx); rm(list = key, envir = .nngp_b200_cache) }
  invisible(TRUE)
}

# GpGp::vecchia_Linv replacement; returns the number of non-positive-definite neighbour blocks
nngp_factor_build = function(ctx, covparms, slot = 0L)
  nngp_check(.C("nngp_factor_build", ctx_id = as.integer(ctx), slot = as.integer(slot), covparms = as.double(covparms),
                n_covparms = length(covparms), n_not_pd = integer(1), status = integer(1)))$n_not_pd

nngp_factor_get = function(ctx, n, m, slot = 0L)
  matrix(nngp_check(.C("nngp_factor_get", ctx_id = as.integer(ctx), slot = as.integer(slot), Linv = double(n * (m + 1)), status = integer(1)))$Linv, n, m + 1)

nngp_field_set = function(ctx, field) invisible(nngp_check(.C("nngp_field_set", ctx_id = as.integer(ctx), field = as.double(field), status = integer(1))))
nngp_field_get = function(ctx, n) nngp_check(.C("nngp_field_get", ctx_id = as.integer(ctx), field = double(n), status = integer(1)))$field
nngp_obs_set = function(ctx, y_minus_xb) invisible(nngp_check(.C("nngp_obs_set", ctx_id = as.integer(ctx), y_minus_xb = as.double(y_minus_xb), status = integer(1))))

# ll_compressed_sparse_chol(Linv, field - beta_0, NNarray, log_scale) on the device-resident field
nngp_loglik = function(ctx, beta_0, log_scale, slot = 0L)
  nngp_check(.C("nngp_loglik", ctx_id = as.integer(ctx), slot = as.integer(slot), beta_0 = as.double(beta_0), log_scale = as.double(log_scale), ll = double(1), status = integer(1)))$ll

# n_sweeps chromatic sweeps (update_Gaussian.R:257-275); z = NULL uses the on-device Philox generator
nngp_gibbs_sweep = function(ctx, n_sweeps, beta_0, log_scale, log_noise_variance, z = NULL, seed = 0)
{
  rng_mode = if(is.null(z)) 1L else 0L
  if(is.null(z)) z = 0
  invisible(nngp_check(.C("nngp_gibbs_sweep", ctx_id = as.integer(ctx), n_sweeps = as.integer(n_sweeps), beta_0 = as.double(beta_0),
                          log_scale = as.double(log_scale), log_noise_variance = as.double(log_noise_variance), rng_mode = rng_mode,
                          z = as.double(z), seed = as.double(seed), status = integer(1))))
}

# one step with state$params$field kept in R: n_sweeps sweeps of `field`, returns list(field = new field, ll = its Vecchia log-likelihood)
nngp_sweep_loglik_host = function(ctx, field, n_sweeps, beta_0, log_scale, log_noise_variance, z = NULL, seed = 0)
{
  rng_mode = if(is.null(z)) 1L else 0L
  if(is.null(z)) z = 0
  res = nngp_check(.C("nngp_sweep_loglik_host", ctx_id = as.integer(ctx), n_sweeps = as.integer(n_sweeps), beta_0 = as.double(beta_0),
                      log_scale = as.double(log_scale), log_noise_variance = as.double(log_noise_variance), rng_mode = rng_mode,
                      z = as.double(z), seed = as.double(seed), field_io = as.double(field), ll = double(1), status = integer(1)))
  list(field = res$field_io, ll = res$ll)
}

# initial field draw (initialize.R:201-208) with R's own rnorm stream
nngp_field_init = function(ctx, beta_0, log_scale, z, slot = 0L)
  invisible(nngp_check(.C("nngp_field_init", ctx_id = as.integer(ctx), slot = as.integer(slot), beta_0 = as.double(beta_0), log_scale = as.double(log_scale), z = as.double(z), status = integer(1))))

# the whole per-chain loop of mcmc_nngp_update_Gaussian (no regressors) behind one call
nngp_chain_run = function(ctx, params, transition_kernels, n_iter, field_thinning, n_chromatic, iter_start, chain_index, var_y, n_locs, rng_mode = 1L)
{
  k = length(params$shape)
  p = c(params$beta_0, params$log_scale, params$log_noise_variance, transition_kernels$covariance_params_sufficient$logvar,
        transition_kernels$covariance_params_ancillary$logvar, params$shape)
  n_frec = round(n_iter * field_thinning)
  nngp_check(.C("nngp_chain_run", ctx_id = as.integer(ctx), n_shape = as.integer(k), params_io = as.double(p), n_iter = as.integer(n_iter),
                thin = as.double(field_thinning), n_chromatic = as.integer(n_chromatic), iter_start = as.integer(iter_start),
                chain_index = as.integer(chain_index), rng_mode = as.integer(rng_mode), var_y = as.double(var_y),
                records_out = double(n_iter * (3 + k)), field_records_out = double(max(n_frec, 1) * n_locs), accept_out = integer(2 * n_iter),
                status = integer(1)))
}

# regressors of the Gaussian model (X as prepared by mcmc_nngp_initialize.R:116-137) uploaded once per context
nngp_regressors_set = function(ctx, X, observed_field, vecchia_approx)
{
  xl = as.integer(X$locs)
  invisible(nngp_check(.C("nngp_regressors_set", ctx_id = as.integer(ctx), p = as.integer(ncol(X$X)), X = as.double(X$X),
                          observed_field = as.double(observed_field), n_xlocs = length(xl), xlocs = if(length(xl)) xl else integer(1),
                          first_obs = as.integer(vecchia_approx$hctam_scol_1), status = integer(1))))
}

# the whole per-chain loop with the regression block (update_Gaussian.R:226-250) behind one call
nngp_chain_run_regressors = function(ctx, params, transition_kernels, X, n_iter, field_thinning, n_chromatic, iter_start, chain_index, var_y,
                                     n_locs, rng_mode = 1L)
{
  k = length(params$shape)
  p = c(params$beta_0, params$log_scale, params$log_noise_variance, transition_kernels$covariance_params_sufficient$logvar,
        transition_kernels$covariance_params_ancillary$logvar, params$shape)
  n_frec = round(n_iter * field_thinning)
  nngp_check(.C("nngp_chain_run_regressors", ctx_id = as.integer(ctx), n_shape = as.integer(k), params_io = as.double(p),
                beta_io = as.double(params$beta), solve_1XT1X = as.double(X$solve_1XT1X), chol_solve_1XT1X = as.double(X$chol_solve_1XT1X),
                n_iter = as.integer(n_iter), thin = as.double(field_thinning), n_chromatic = as.integer(n_chromatic),
                iter_start = as.integer(iter_start), chain_index = as.integer(chain_index), rng_mode = as.integer(rng_mode),
                var_y = as.double(var_y), records_out = double(n_iter * (3 + k)), beta_records_out = double(n_iter * ncol(X$X)),
                field_records_out = double(max(n_frec, 1) * n_locs), accept_out = integer(2 * n_iter), status = integer(1)))
}

# all chains of a cycle behind ONE call, advanced concurrently (mclapply over chains, update_Gaussian.R:22-26): chain k runs on
# ctxs[k] (its own GPU / stream / host thread), at most max_concurrent (= n_cores) in flight.  params: n_chains x (5 + k) matrix
# with ROWS [beta_0, log_scale, log_noise_variance, logvar_sufficient, logvar_ancillary, shape...]; per-chain blocks come back
# chain after chain.  .C() cannot carry long vectors: n_chains * round(n_iter * thin) * n_locs must stay below 2^31 (lower
# field_thinning or use the .Call glue of R/r_glue.c, which has no such limit).
nngp_chains_run = function(ctxs, params, n_iter, field_thinning, n_chromatic, iter_start, chain_index, var_y, n_locs, rng_mode = 1L,
                           max_concurrent = length(ctxs))
{
  nc = length(ctxs); k = ncol(params) - 5L
  n_frec = round(n_iter * field_thinning)
  if(as.double(nc) * max(n_frec, 1) * n_locs >= 2^31) stop("field records exceed what the dot-C interface can carry (2^31 - 1 doubles): lower field_thinning or use R/r_glue.c")
  nngp_check(.C("nngp_chains_run", n_chains = as.integer(nc), ctx_ids = as.integer(ctxs), n_shape = as.integer(k), params_io = as.double(t(params)),
                n_iter = as.integer(n_iter), thin = as.double(field_thinning), n_chromatic = as.integer(n_chromatic),
                iter_start = as.integer(iter_start), chain_index = as.integer(chain_index), rng_mode = as.integer(rng_mode),
                var_y = as.double(var_y), max_concurrent = as.integer(max(1L, max_concurrent)), records_out = double(nc * n_iter * (3 + k)),
                field_records_out = double(nc * max(n_frec, 1) * n_locs), accept_out = integer(nc * 2 * n_iter), status = integer(1)))
}

nngp_chains_run_regressors = function(ctxs, params, beta, X, n_iter, field_thinning, n_chromatic, iter_start, chain_index, var_y, n_locs,
                                      rng_mode = 1L, max_concurrent = length(ctxs))
{
  nc = length(ctxs); k = ncol(params) - 5L; p = ncol(X$X)
  n_frec = round(n_iter * field_thinning)
  if(as.double(nc) * max(n_frec, 1) * n_locs >= 2^31) stop("field records exceed what the dot-C interface can carry (2^31 - 1 doubles): lower field_thinning or use R/r_glue.c")
  nngp_check(.C("nngp_chains_run_regressors", n_chains = as.integer(nc), ctx_ids = as.integer(ctxs), n_shape = as.integer(k),
                params_io = as.double(t(params)), beta_io = as.double(t(beta)), solve_1XT1X = as.double(X$solve_1XT1X),
                chol_solve_1XT1X = as.double(X$chol_solve_1XT1X), n_iter = as.integer(n_iter), thin = as.double(field_thinning),
                n_chromatic = as.integer(n_chromatic), iter_start = as.integer(iter_start), chain_index = as.integer(chain_index),
                rng_mode = as.integer(rng_mode), var_y = as.double(var_y), max_concurrent = as.integer(max(1L, max_concurrent)),
                records_out = double(nc * n_iter * (3 + k)), beta_records_out = double(nc * n_iter * p),
                field_records_out = double(nc * max(n_frec, 1) * n_locs), accept_out = integer(nc * 2 * n_iter), status = integer(1)))
}

# GpGp::find_ordered_nn replacement (exact m nearest previous sites, ties by lower index; initialize.R:93, predict.R:5)
nngp_find_ordered_nn = function(locs, m)
{
  locs = as.matrix(locs)
  res = nngp_check(.C("nngp_host_find_ordered_nn", locs = as.double(locs), n = nrow(locs), d = ncol(locs), m = as.integer(m),
                      NNarray = integer(nrow(locs) * (m + 1)), status = integer(1), NAOK = TRUE))
  matrix(res$NNarray, nrow(locs), m + 1)   # INT_MIN comes back as NA_integer_, exactly like GpGp's NA slots
}

# context over the joint (observed ++ predicted) site set of mcmc_nngp_predict_field: no colouring (all zero), no observations
nngp_ctx_create_predict = function(locs, NNarray, stationary_covfun, device = 0L, layout = 2L)
{
  locs = as.matrix(locs)
  res = nngp_check(.C("nngp_ctx_create", n = nrow(locs), d = ncol(locs), m = as.integer(ncol(NNarray) - 1L), locs = as.double(locs),
                      NNarray = as.integer(NNarray), coloring = integer(nrow(locs)), n_obs = 0L, locs_match = integer(1),
                      covfun_id = nngp_covfun_id(stationary_covfun), device = as.integer(device), layout = as.integer(layout),
                      ctx_id = integer(1), status = integer(1), NAOK = TRUE))
  res$ctx_id
}

# one stored sample conditionally simulated at the new sites (predict.R:43-53); z = rnorm(n_pred)
nngp_predict_sample = function(ctx, n_obs_sites, field, beta_0, log_scale, z, slot = 0L)
  nngp_check(.C("nngp_predict_sample", ctx_id = as.integer(ctx), slot = as.integer(slot), n_obs_sites = as.integer(n_obs_sites),
                field = as.double(field), beta_0 = as.double(beta_0), log_scale = as.double(log_scale), z_pred = as.double(z),
                out = double(length(z)), status = integer(1)))$out

# Replaces the crossprod() moral graph + naive_greedy_coloring of mcmc_nngp_initialize.R:103-110 / Coloring.R:2-20 (the R loop
# needs an (n+1) x maxdeg double scratch: 2.2 GB at n = 1M, ~47 GB at n = 10M): identical colours, O(nnz) memory.
# In mcmc_nngp_initialize:   vecchia_approx$coloring = nngp_greedy_coloring(vecchia_approx$NNarray)
nngp_greedy_coloring = function(NNarray)
  nngp_check(.C("nngp_host_greedy_coloring", NNarray = as.integer(NNarray), n = nrow(NNarray), m = as.integer(ncol(NNarray) - 1L),
                coloring = integer(nrow(NNarray)), n_colors = integer(1), status = integer(1), NAOK = TRUE))$coloring

# exact farthest-point ("max-min") ordering; GpGp::order_maxmin (initialize.R:29) is a randomised approximation of it, so the
# two orderings differ -- any ordering is valid for the model, and an ordering computed by GpGp can be used unchanged
nngp_order_maxmin = function(locs)
{
  locs = as.matrix(locs)
  nngp_check(.C("nngp_host_order_maxmin", locs = as.double(locs), n = nrow(locs), d = ncol(locs), order = integer(nrow(locs)), status = integer(1)))$order
}

# the same colouring from the MRF adjacency matrix itself (dgCMatrix slots): what R/Coloring.R, the drop-in for
# Scripts/Coloring.R, calls
nngp_greedy_coloring_adj = function(M)
  nngp_check(.C("nngp_host_greedy_coloring_adj", adj_p = as.integer(M@p), adj_i = as.integer(M@i), n = nrow(M), coloring = integer(nrow(M)),
                n_colors = integer(1), status = integer(1)))$coloring
