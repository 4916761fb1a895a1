#!/bin/bash
# one bench line per BASELINE-shaped workload -> gpurun_out/w_<name>_<mode>.json
mkdir -p gpurun_out
run() { name=$1; shift; python bench.py "$@" --steps 10 --warmup 10 --no-cpu-baseline > gpurun_out/w_$name.json 2>/dev/null; echo "$name rc=$?"; }
run family_eval --workload family
run family_train --workload family --train --batch 20
run fb237v2_eval --workload fb237v2
run fb237v2_train --workload fb237v2 --train --batch 10
run fb15k237_train --workload fb15k237 --train --batch 16
run yago310_eval --workload yago310
run yago310_train --workload yago310 --train --batch 4
run powerlaw_eval --workload powerlaw
run powerlaw_train --workload powerlaw --train --batch 2
