#!/bin/bash
# sweep REDGNN_GRAD_COPIES on the training bench
for c in "$@"; do
  REDGNN_GRAD_COPIES=$c python bench.py --train --batch 16 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('copies $c', round(d['ms_per_step'],3), d['kernel_ms_per_step'])"
done
