#!/bin/bash
# the default bench line and the reference arm, as the driver runs them
TAG=${1:-bench}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
tail -3 $OUT/bench_default.err
python - <<PY
import json
d=json.load(open("$OUT/bench_default.json"))
def show(o):
    r=o.get("roofline") or {}; v=o.get("verified") or {}; e=o.get("e2e") or {}
    print(o["config"]["workload"], "ms %.4f"%o.get("ms_per_step",0), "call %.4f"%(r.get("avg_launch_ms") or 0), "frac %.3f"%r.get("frac",0), "e2e %.3g"%(e.get("value") or 0), "verified", v.get("ok"), o.get("error"))
show(d)
for o in d.get("other_workloads",[]): show(o)
print(d.get("other_workloads_skipped"), d["clocks"])
PY
