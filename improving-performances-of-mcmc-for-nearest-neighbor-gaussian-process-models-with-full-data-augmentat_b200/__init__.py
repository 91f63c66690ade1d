"""nngp_b200 -- B200-native hot path of the NNGP full-data-augmentation sampler.

Host-side mirror of the reference's R entry points (mcmc_nngp_initialize / mcmc_nngp_run / mcmc_nngp_update_Gaussian /
mcmc_nngp_predict_field / mcmc_nngp_estimate) over the C ABI of libnngp_b200.so (include/nngp_b200.h).  R is not available
in this image, so the host side above the ABI is Python; R/ holds the equivalent .C() glue.
"""
from . import _lib
from ._lib import (COVFUN_IDS, LAYOUT_COLOR, LAYOUT_COLOR_MORTON, LAYOUT_MORTON, NA_INT, RNG_PHILOX, RNG_SUPPLIED, SLOT_CURRENT,
                   SLOT_PROPOSAL, NNGPError, PinnedArray, device_count, launch_count, set_host_threads)
from .context import NNGPContext, chains_run, find_ordered_nn, fp64_peak, greedy_coloring, order_maxmin, time_op_group
from .partition import shard_plan, spatial_blocks
from .rstream import RStream, find_ordered_nn_gpgp, order_maxmin_gpgp
from .sharded import (ShardedContext, comm_unique_id, connect_local, create_sharded_distributed, group_chain_run, group_loglik, group_sweep,
                      host_routed_sweep)
from .api import (ESS, Gelman_Rubin_Brooks, get_summary, mcmc_nngp_estimate, mcmc_nngp_initialize, mcmc_nngp_predict,
                  mcmc_nngp_predict_field, mcmc_nngp_predict_fixed_effects, mcmc_nngp_run, mcmc_nngp_update_Gaussian,
                  release_contexts)

__all__ = ["mcmc_nngp_initialize", "mcmc_nngp_run", "mcmc_nngp_update_Gaussian", "mcmc_nngp_predict_field", "mcmc_nngp_predict",
           "mcmc_nngp_predict_fixed_effects", "mcmc_nngp_estimate", "get_summary", "Gelman_Rubin_Brooks", "ESS", "release_contexts",
           "ShardedContext", "shard_plan", "spatial_blocks", "host_routed_sweep", "connect_local", "group_sweep", "group_loglik", "group_chain_run", "create_sharded_distributed", "comm_unique_id",
           "NNGPContext", "chains_run", "time_op_group", "fp64_peak", "find_ordered_nn", "greedy_coloring", "order_maxmin", "RStream", "order_maxmin_gpgp", "find_ordered_nn_gpgp", "NNGPError", "PinnedArray", "device_count", "launch_count", "set_host_threads",
           "COVFUN_IDS", "NA_INT", "SLOT_CURRENT", "SLOT_PROPOSAL", "RNG_SUPPLIED", "RNG_PHILOX", "LAYOUT_COLOR",
           "LAYOUT_COLOR_MORTON", "LAYOUT_MORTON"]
