# R/mcmc_nngp_update_Gaussian.R -- drop-in for Scripts/mcmc_nngp_update_Gaussian.R of the reference: same signature, same
# list(state, records) per chain, but the iteration loop (reference lines 101-314) runs on the GPU behind nngp_chain_run.
# Chains are NOT forked (a CUDA context does not survive fork(); reference line 25 uses mclapply): they are dispatched
# round-robin over the visible GPUs inside this process.  Untested in the build image (no R there); see INTEGRATION.md.
source(file.path("R", "nngp_b200.R"))

mcmc_nngp_update_Gaussian = function(locs, X, observed_field, space_time_model, vecchia_approx, states, n_iterations_update,
                                     n_cores = NULL, field_thinning = 1, ancillary = T, n_chromatic = 10, iterations,
                                     n_gpus = 1, rng = c("philox", "R"))
{
  nngp_b200_load()
  rng_mode = if(match.arg(rng) == "R") 0L else 1L
  iter_start = iterations[nrow(iterations), 1]
  n_locs = vecchia_approx$n_locs
  k = length(space_time_model$covfun$shape_params)
  lapply(seq(length(states)), function(i)
  {
    state = states[[i]]
    ctx = nngp_ctx_create(locs, vecchia_approx, space_time_model$covfun$stationary_covfun, device = (i - 1L) %% n_gpus)
    on.exit(nngp_ctx_destroy(ctx))
    nngp_field_set(ctx, state$params$field)
    if(is.null(X$X))
    {
      nngp_obs_set(ctx, observed_field)   # mu - beta_0 = 0 without regressors
      res = nngp_chain_run(ctx, state$params, state$transition_kernels, n_iterations_update, field_thinning, n_chromatic, iter_start, i,
                           var(observed_field), n_locs, rng_mode)
    }
    else   # reference lines 226-250 on the device: X$X stays in HBM, only (p+1)-vectors come back per iteration
    {
      nngp_regressors_set(ctx, X, observed_field, vecchia_approx)
      res = nngp_chain_run_regressors(ctx, state$params, state$transition_kernels, X, n_iterations_update, field_thinning, n_chromatic,
                                      iter_start, i, var(observed_field), n_locs, rng_mode)
      beta_names = names(state$params$beta)
      state$params$beta = res$beta_io; names(state$params$beta) = beta_names
    }
    p = res$params_io
    state$params$beta_0 = p[1]; state$params$log_scale = p[2]; state$params$log_noise_variance = p[3]
    state$transition_kernels$covariance_params_sufficient$logvar = p[4]
    state$transition_kernels$covariance_params_ancillary$logvar = p[5]
    state$params$shape = p[5 + seq(k)]
    state$params$field = nngp_field_get(ctx, n_locs)
    rec = matrix(res$records_out, n_iterations_update, 3 + k)
    records = list()
    records$beta_0 = matrix(rec[, 1], ncol = 1); colnames(records$beta_0) = "beta_0"
    if(!is.null(X$X))
    {
      records$beta = matrix(res$beta_records_out, n_iterations_update, ncol(X$X)); colnames(records$beta) = names(state$params$beta)
    }
    records$log_scale = matrix(rec[, 2], ncol = 1)
    records$log_noise_variance = matrix(rec[, 3], ncol = 1)
    records$shape = matrix(rec[, 3 + seq(k)], ncol = k); colnames(records$shape) = space_time_model$covfun$shape_params
    records$field = matrix(res$field_records_out[seq(round(n_iterations_update * field_thinning) * n_locs)], ncol = n_locs)
    list("state" = state, "records" = records)
  })
}
