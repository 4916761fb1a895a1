#!/bin/bash
# eval forward on three workloads (GPU ms/step from the CUPTI breakdown)
for w in "fb15k237 64" "yago310 8" "powerlaw 4"; do python scratch/profile_eval.py $w 2>&1 | grep -E "workload|k_edge_fwd" ; done
