# R/mcmc_nngp_update_Gaussian.R -- drop-in for Scripts/mcmc_nngp_update_Gaussian.R of the reference: same signature, same
# list(state, records) per chain, but the iteration loop (reference lines 101-314) runs on the GPU.  All chains of the cycle
# advance CONCURRENTLY behind one nngp_chains_run call -- chain i on GPU (i - 1) mod n_gpus, its own stream and host thread, at
# most n_cores in flight -- which is what the reference's mclapply (line 25) does with processes (a CUDA context does not
# survive fork()).  Contexts are cached between cycles (nngp_chain_context); nngp_b200_release() frees them.
# Untested in the build image (no R there); see INTEGRATION.md.
if(!exists("nngp_b200_load")) source(file.path(Sys.getenv("NNGP_B200_HOME", unset = "."), "R", "nngp_b200.R"))

mcmc_nngp_update_Gaussian = function(locs, X, observed_field, space_time_model, vecchia_approx, states, n_iterations_update,
                                     n_cores = NULL, field_thinning = 1, ancillary = T, n_chromatic = 10, iterations,
                                     n_gpus = getOption("nngp_b200.n_gpus", 1L), rng = getOption("nngp_b200.rng", "philox"))
{
  # mcmc_nngp_run calls this function with the reference's arguments only, so the two extra ones default to options:
  #   options(nngp_b200.n_gpus = 8L)    chain i on GPU (i - 1) mod n_gpus
  #   options(nngp_b200.rng = "R")      R's own stream, seeded iter_start + i as in the reference (line 36): the chains then draw
  #                                     exactly what they draw in the reference; "philox" (default) = the on-device generator
  nngp_b200_load()
  if(!rng %in% c("philox", "R")) stop(paste("unknown rng", rng))
  rng_mode = if(rng == "R") 0L else 1L
  iter_start = iterations[nrow(iterations), 1]
  n_locs = vecchia_approx$n_locs
  n_chains = length(states)
  k = length(space_time_model$covfun$shape_params)
  if(is.null(n_cores)) n_cores = n_chains
  covfun = space_time_model$covfun$stationary_covfun
  ctxs = sapply(seq_len(n_chains), function(i) nngp_chain_context(i, locs, vecchia_approx, covfun, device = (i - 1L) %% n_gpus))
  params = t(sapply(states, function(s) c(s$params$beta_0, s$params$log_scale, s$params$log_noise_variance,
                                          s$transition_kernels$covariance_params_sufficient$logvar,
                                          s$transition_kernels$covariance_params_ancillary$logvar, s$params$shape)))
  for(i in seq_len(n_chains))
  {
    nngp_field_set(ctxs[i], states[[i]]$params$field)
    if(is.null(X$X)) nngp_obs_set(ctxs[i], observed_field)   # mu - beta_0 = 0 without regressors
    else if(!.nngp_b200_cache[[paste0("chain_", i)]]$regressors)
    {
      nngp_regressors_set(ctxs[i], X, observed_field, vecchia_approx)   # X$X stays in HBM for every later cycle
      .nngp_b200_cache[[paste0("chain_", i)]]$regressors = TRUE
    }
  }
  if(is.null(X$X))
    res = nngp_chains_run(ctxs, params, n_iterations_update, field_thinning, n_chromatic, iter_start, seq_len(n_chains), var(observed_field),
                          n_locs, rng_mode, n_cores)
  else   # reference lines 226-250 on the device: only (p+1)-vectors come back per iteration
    res = nngp_chains_run_regressors(ctxs, params, t(sapply(states, function(s) s$params$beta)), X, n_iterations_update, field_thinning,
                                     n_chromatic, iter_start, seq_len(n_chains), var(observed_field), n_locs, rng_mode, n_cores)
  p_out = matrix(res$params_io, nrow = n_chains, byrow = TRUE)
  n_frec = round(n_iterations_update * field_thinning)
  lapply(seq_len(n_chains), function(i)
  {
    state = states[[i]]
    p = p_out[i, ]
    state$params$beta_0 = p[1]; state$params$log_scale = p[2]; state$params$log_noise_variance = p[3]
    state$transition_kernels$covariance_params_sufficient$logvar = p[4]
    state$transition_kernels$covariance_params_ancillary$logvar = p[5]
    state$params$shape = p[5 + seq_len(k)]
    state$params$field = nngp_field_get(ctxs[i], n_locs)
    rec = matrix(res$records_out[(i - 1) * n_iterations_update * (3 + k) + seq_len(n_iterations_update * (3 + k))], n_iterations_update, 3 + k)
    records = list()
    records$beta_0 = matrix(rec[, 1], ncol = 1); colnames(records$beta_0) = "beta_0"
    if(!is.null(X$X))
    {
      pb = ncol(X$X)
      beta_names = names(state$params$beta)
      state$params$beta = res$beta_io[(i - 1) * pb + seq_len(pb)]; names(state$params$beta) = beta_names
      records$beta = matrix(res$beta_records_out[(i - 1) * n_iterations_update * pb + seq_len(n_iterations_update * pb)], n_iterations_update, pb)
      colnames(records$beta) = beta_names
    }
    records$log_scale = matrix(rec[, 2], ncol = 1)
    records$log_noise_variance = matrix(rec[, 3], ncol = 1)
    records$shape = matrix(rec[, 3 + seq_len(k)], ncol = k); colnames(records$shape) = space_time_model$covfun$shape_params
    # n_frec x n_locs; zero rows when round(n_iter * thinning) == 0, like the reference's records$field
    records$field = matrix(res$field_records_out[(i - 1) * max(n_frec, 1) * n_locs + seq_len(n_frec * n_locs)], nrow = n_frec, ncol = n_locs)
    list("state" = state, "records" = records)
  })
}
