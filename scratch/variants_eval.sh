#!/bin/bash
# eval bench with each scratch/exp/lib_<name>.so.  usage: variants_eval.sh "<bench args>" name...
args=$1; shift
for v in "$@"; do
  REDGNN_B200_LIB=$PWD/scratch/exp/lib_$v.so python bench.py $args --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', round(d['ms_per_step'],3), d['kernel_ms_per_step'])"
done
