#!/bin/bash
OUT=gpurun_out/r2f
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for i in 1 2; do
( time timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_default_$i.json 2> $OUT/bench_default_$i.err ) 2>&1 | grep real
grep -i "fail\|error" $OUT/bench_default_$i.err | head
python - <<PY
import json
d=json.load(open("$OUT/bench_default_$i.json"))
def show(o):
    r=o.get("roofline") or {}
    v=o.get("verified") or {}
    print(o["config"]["workload"], "ms %.3f"%o.get("ms_per_step",0), "frac %.3f"%r.get("frac",0), "e2e %.3g"%((o.get("e2e") or {}).get("value") or 0), "cpu %.3g"%((o.get("cpu_baseline") or {}).get("value") or 0), "verified", v.get("ok"), v.get("pixels_differing_from_oracle"), v.get("first_differences"), o.get("error"))
show(d)
for o in d.get("other_workloads",[]): show(o)
PY
done
