mkdir -p gpurun_out; rm -f gpurun_out/r15_*
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:"bn_|stem_|pack" --csv --log-file gpurun_out/r15_bn.csv python scripts/prof_step.py 256 > gpurun_out/r15_bn.log 2>&1
tail -3 gpurun_out/r15_bn.log
