set -u
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,sm__cycles_active.avg"
for n in 1000000 4000000; do
 for v in 6 22; do
  ncu --metrics $M --clock-control none -k regex:gibbs_tile --launch-skip 44 --launch-count 22 --csv --log-file gpurun_out/ncu_cmp_${n}_v${v}.csv python scripts/sweep_once.py $n $v 3 > gpurun_out/ncu_cmp_${n}_v${v}.log 2>&1
  echo "n=$n v=$v rc=$?"
 done
done
