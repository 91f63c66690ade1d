#!/bin/bash
# ncu launch list of the bench command + one full capture of the factor kernel (run under gpurun, one GPU)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-predict"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
rm -f gpurun_out/*.ncu-rep gpurun_out/prof_*_raw.csv gpurun_out/prof_*_source.csv
ncu --set full --clock-control none --import-source on -k regex:vecchia_factor_reg_kernel -s 1 -c 1 -f -o gpurun_out/prof_factor $CMD > gpurun_out/ncu_factor.log 2>&1
echo "factor full rc=$?"
ncu -i gpurun_out/prof_factor.ncu-rep --page raw --csv > gpurun_out/prof_factor_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_factor.ncu-rep --page source --csv > gpurun_out/prof_factor_source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls gpurun_out
