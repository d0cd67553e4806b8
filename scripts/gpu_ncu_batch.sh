#!/bin/bash
# ncu captures of the solve kernel inside bench.py's own command line (outputs digested on the box: gpurun_out/ is capped at 64 MiB)
#   FULL="2:4096"            configs:batches captured with --set full --import-source on (report kept)
#   LIGHT="2:32768 4:0"      configs:batches captured with the traffic / issue metrics only (batch 0 = the config's own)
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct,sm__inst_executed.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__cycles_elapsed.max
for spec in ${FULL:-2:4096}; do
  cfg=${spec%%:*}; b=${spec##*:}; name=r2_cfg${cfg}_b${b}
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:nmpc_solve_kernel -s 4 -c 1 -f -o gpurun_out/$name python bench.py --config $cfg --batch $b --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv
  ncu -i gpurun_out/$name.ncu-rep --page source --csv --print-source cuda > gpurun_out/${name}_source.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt
  ls -la gpurun_out/$name*
done
for spec in $LIGHT; do
  cfg=${spec%%:*}; b=${spec##*:}; name=r2_cfg${cfg}_b${b}
  timeout 900 ncu --metrics $M --clock-control none -k regex:nmpc_solve_kernel -s 4 -c 1 --csv --log-file gpurun_out/${name}_metrics.csv python bench.py --config $cfg --batch $b --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_$name.log 2>&1
  tail -16 gpurun_out/${name}_metrics.csv | cut -c1-200
done
if [ -n "$LAUNCHES" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
  tail -5 gpurun_out/r2_launches_cfg2.csv | cut -c1-250
fi
