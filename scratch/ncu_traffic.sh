#!/bin/bash
# DRAM traffic / L2 hit / L1 wavefronts of every fused edge-forward launch of one bench run per workload
out=gpurun_out/${1:-ncu}; mkdir -p $out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for w in fb15k237 yago310 powerlaw; do
  timeout 900 ncu --metrics $M --clock-control none -k regex:"k_edge_fwd" --csv --log-file $out/edge_fwd_$w.csv \
      python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-train-subsystem > $out/edge_fwd_$w.log 2>&1
  echo "$w ncu rc=$?"
done
