#!/bin/bash
# bench.py on N GPUs of one box (torchrun): FB15k-237 default line (eval + train subsystem), power-law training,
# and (N = 8) the YAGO3-10-shaped eval; JSON lines under gpurun_out/$2/.
N=$1; out=gpurun_out/${2:-multi}; mkdir -p $out
run() { name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --no-cpu-baseline "$@" > $out/${name}_${N}gpu.json 2> $out/${name}_${N}gpu.err; echo "$name rc=$?"; }
run fb15k237
run powerlaw_train --workload powerlaw --train
[ "$N" = 8 ] && run yago310 --workload yago310
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/*_${N}gpu.json")):
    d = json.load(open(f)); t = d["subsystems"]["train"] or {}
    print(f.split("/")[-1], "n", d["n_gpus"], "ms", round(d["ms_per_step"], 3), "q/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
          "| train ms", round(t.get("ms_per_step", 0), 3), "q/s", round(t.get("queries_per_s", 0), 1),
          "allreduce us", (t.get("allreduce") or {}).get("us_per_step"), (t.get("dist_check") or {}).get("status"), (t.get("dist_check") or {}).get("max_rel_err"))
PY
