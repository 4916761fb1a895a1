#!/bin/bash
for v in "$@"; do
  echo "== variant $v"
  export REDGNN_B200_LIB=$PWD/scratch/exp/lib_$v.so
  for w in "fb15k237 64" "yago310 8" "powerlaw 4"; do python scratch/profile_eval.py $w 2>&1 | grep -E "workload|k_edge_fwd(_p)?<48, true" ; done
done
