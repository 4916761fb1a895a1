#!/bin/bash
# ncu captures behind profiles/: (1) DRAM traffic / L2 hit / L1 wavefronts of every fused edge-forward launch of
# one bench run per workload (light metric set), (2) --set full of the training kernels of one FB15k-237 step.
out=gpurun_out/${1:-ncu}; mkdir -p $out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for w in fb15k237 yago310 powerlaw; do
  timeout 900 ncu --metrics $M --clock-control none -k regex:"k_edge_fwd" --csv --log-file $out/edge_fwd_$w.csv \
      python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-train-subsystem > $out/edge_fwd_$w.log 2>&1
  echo "$w ncu rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_node_bwd_tc|k_node_wgrad|k_edge_bwd_p|k_node_update_tc|k_edge_fwd_p|k_attn_param_grads" -s 56 -c 28 \
    -o $out/train_kernels python scratch/train_step.py 16 3 > $out/train_kernels.log 2>&1; echo "train ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $out/launches_train.csv \
    python scratch/train_step.py 16 2 > $out/launches_train.log 2>&1; echo "launch list train rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/launches_eval.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-subsystem > $out/launches_eval.log 2>&1; echo "launch list eval rc=$?"
