# R/mcmc_nngp_predict.R -- drop-in for mcmc_nngp_predict_field of the reference (Scripts/mcmc_nngp_predict.R:1-60): same
# arguments, same list(predicted_locs, predicted_field_samples, predicted_field_summary).  The joint neighbour table is built
# by the library's host utility, the Vecchia factor of the joint site set by nngp_factor_build (only when the shape
# parameters differ from the previous stored sample's; the reference factors once per distinct shape row via !duplicated(),
# predict.R:24 -- the same set of factorisations, since a shape row only repeats on consecutive rejected iterations), and each
# stored field sample is carried to the new sites
# by nngp_predict_sample, which solves the new rows only.  Chains are processed one after the other on one GPU (the
# reference forks them, :16; a CUDA context does not survive fork()).  Untested in the build image (no R there).
if(!exists("nngp_b200_load")) source(file.path(Sys.getenv("NNGP_B200_HOME", unset = "."), "R", "nngp_b200.R"))

mcmc_nngp_predict_field = function(mcmc_nngp_list, predicted_locs, burn_in = .5, n_cores = 1, m = 10, device = 0L)
{
  nngp_b200_load()
  predicted_locs = as.matrix(predicted_locs)
  joint_locs = rbind(mcmc_nngp_list$locs, predicted_locs)
  n_locs = mcmc_nngp_list$vecchia_approx$n_locs
  n_pred = nrow(predicted_locs)
  covfun = mcmc_nngp_list$space_time_model$covfun
  ctx = nngp_ctx_create_predict(joint_locs, nngp_find_ordered_nn(joint_locs, m), covfun$stationary_covfun, device)
  on.exit(nngp_ctx_destroy(ctx))
  kept = mcmc_nngp_list$records$chain_1$saved_field
  kept = kept[kept > burn_in * max(kept)]
  # reference quirk kept on purpose: here the "qlogis" smoothness is mapped with 1.5 * plogis (predict.R:37), not .5 + .5 * plogis
  to_covparms = function(shape_row) c(1, mapply(function(name, v) if(substr(name, 1, 3) == "log") exp(v) else 1.5 * plogis(v),
                                                covfun$shape_params, shape_row), 0)
  samples = lapply(mcmc_nngp_list$records, function(chain)
  {
    out = matrix(0, length(kept), n_pred)
    last_shape = NULL
    for(k in seq_along(kept))
    {
      it = kept[k]
      shape_row = chain$params$shape[it, ]
      if(is.null(last_shape) || any(shape_row != last_shape))
      {
        nngp_factor_build(ctx, to_covparms(shape_row))
        last_shape = shape_row
      }
      out[k, ] = nngp_predict_sample(ctx, n_locs, chain$params$field[match(it, chain$saved_field), ], chain$params$beta_0[it],
                                     chain$params$log_scale[it], rnorm(n_pred))
    }
    out
  })
  list("predicted_locs" = predicted_locs, "predicted_field_samples" = samples,
       "predicted_field_summary" = get_summary(do.call(rbind, samples)))
}
