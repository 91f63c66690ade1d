#!/bin/bash
# Short ncu pass (run under gpurun, one GPU): launch list of the bench command + the sweep kernel (22 colour launches with
# the memory sections, one full capture of the first colour) + the new regression kernels.  See profile_run.sh for the long one.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-predict"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
rm -f gpurun_out/*.ncu-rep
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats --clock-control none --cache-control none -k regex:gibbs_tile2_kernel -s 44 -c 22 -f -o gpurun_out/prof_gibbs $CMD > gpurun_out/ncu_gibbs.log 2>&1
echo "gibbs sections rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gibbs_tile2_kernel -s 44 -c 1 -f -o gpurun_out/prof_gibbs_first $CMD > gpurun_out/ncu_gibbs_first.log 2>&1
echo "gibbs full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"transpose_tile2_kernel|sptrsv_syncfree_kernel" -s 1 -c 2 -f -o gpurun_out/prof_other $CMD > gpurun_out/ncu_other.log 2>&1
echo "other full rc=$?"
for r in gpurun_out/*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > ${b}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/prof_gibbs_first.ncu-rep --page source --csv > gpurun_out/prof_gibbs_first_source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out
