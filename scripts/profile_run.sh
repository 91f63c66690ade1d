#!/bin/bash
# ncu evidence for the hot kernels (run under gpurun, one GPU).  Writes gpurun_out/*.csv / *.ncu-rep.
# The plain run of the same command line must exit 0 first (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-predict --no-multichain --no-other-ordering --no-chain"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# one full sweep = K colour launches of gibbs_tile2_kernel (K = 22 at n = 1M): skip the first two sweeps.
# --cache-control none: the sweep's DRAM traffic is reported warm (L2 as the previous launches left it), like the timed run
rm -f gpurun_out/*.ncu-rep   # the whole directory must stay below 64 MiB to be copied back
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats --clock-control none --cache-control none -k regex:gibbs_tile2_kernel -s 44 -c 22 -f -o gpurun_out/prof_gibbs $CMD > gpurun_out/ncu_gibbs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gibbs_tile2_kernel -s 44 -c 1 -f -o gpurun_out/prof_gibbs_first $CMD > gpurun_out/ncu_gibbs_first.log 2>&1
echo "gibbs full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:vecchia_factor_reg_kernel -s 1 -c 1 -f -o gpurun_out/prof_factor $CMD > gpurun_out/ncu_factor.log 2>&1
echo "factor full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"loglik_partial_kernel" -s 2 -c 1 -f -o gpurun_out/prof_loglik $CMD > gpurun_out/ncu_loglik.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"transpose_tile2_kernel|sptrsv_syncfree_kernel" -s 1 -c 2 -f -o gpurun_out/prof_other $CMD > gpurun_out/ncu_other.log 2>&1
echo "other full rc=$?"
# gpurun copies back at most 64 MiB and every .ncu-rep carries ~14 MB of module image: keep the raw CSV pages instead
for r in gpurun_out/*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > ${b}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/prof_gibbs_first.ncu-rep --page source --csv > gpurun_out/prof_gibbs_first_source.csv 2>/dev/null
ncu -i gpurun_out/prof_factor.ncu-rep --page source --csv > gpurun_out/prof_factor_source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out
