# R/Coloring.R -- drop-in for Scripts/Coloring.R of the reference: same function, same argument (the MRF adjacency matrix built at
# Scripts/mcmc_nngp_initialize.R:103-109), same colours 1..K.  The reference loop (Coloring.R:2-20) allocates an (n+1) x maxdeg
# matrix of doubles (2.2 GB at n = 1M, ~47 GB at n = 10M, m = 20) and runs n interpreted iterations; this one walks the
# compressed-column slots of M once in the library (O(n + nnz)).  Source it INSTEAD of Scripts/Coloring.R; nothing else changes.
# Untested in the build image (no R there); the native routine is checked against the oracle's transcription of Coloring.R.
if(!exists("nngp_b200_load")) source(file.path(Sys.getenv("NNGP_B200_HOME", unset = "."), "R", "nngp_b200.R"))

naive_greedy_coloring = function(M)
{
  nngp_b200_load()
  nngp_greedy_coloring_adj(M)
}
