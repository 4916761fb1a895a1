#!/bin/bash
# Every BASELINE workload through bench.py on one GPU (eval headline + train subsystem, and --train as headline),
# one JSON line each under gpurun_out/$1/.
out=gpurun_out/${1:-workloads}; mkdir -p $out
for w in fb15k237 family fb237v2 yago310 powerlaw; do
  cq=2; [ $w = yago310 ] && cq=1; [ $w = powerlaw ] && cq=1
  timeout 900 python bench.py --workload $w --cpu-queries $cq > $out/${w}_eval.json 2> $out/${w}_eval.err; echo "$w eval rc=$?"
  timeout 900 python bench.py --workload $w --train --cpu-queries 1 > $out/${w}_train.json 2> $out/${w}_train.err; echo "$w train rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    t = d["subsystems"]["train"] or {}
    print(f.split("/")[-1], "ms", round(d["ms_per_step"], 3), "q/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
          "fwd frac", round(d["roofline"]["frac"], 3), "expand", round(d["subsystems"]["expand"]["frac"], 3),
          "train ms", round(t.get("ms_per_step", 0), 3), "bwd frac", round((t.get("edge_bwd") or {}).get("frac", 0), 3),
          "cpu", round(d["cpu_baseline"]["value"], 4), d["cpu_baseline"]["kind"], d["data"][:9])
PY
