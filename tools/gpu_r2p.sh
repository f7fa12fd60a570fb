#!/bin/bash
OUT=gpurun_out/r2p
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/ -x -q -m gpu -s > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt
grep -E "passed|failed|serial .* ms, 6 threads" $OUT/pytest_gpu.txt | tail -5
( time timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2>&1 | grep real
grep -i "fail\|error\|Traceback" -A3 $OUT/bench_default.err | head -20
python - <<PY
import json
d=json.load(open("$OUT/bench_default.json"))
def show(o):
    r=o.get("roofline") or {}
    v=o.get("verified") or {}
    e=o.get("e2e") or {}
    dk=(r.get("dominant_kernel") or {})
    print(o["config"]["workload"], "ms %.3f"%o.get("ms_per_step",0), "call %.3f"%r.get("avg_launch_ms",0), "frac %.3f"%r.get("frac",0), "main %.3f"%dk.get("avg_launch_ms",0), "e2e %.3g"%(e.get("value") or 0), "pageable", (e.get("pageable") or {}).get("h2d_gbs"), "cpu %.3g"%((o.get("cpu_baseline") or {}).get("value") or 0), "verified", v.get("ok"), v.get("pixels_differing_from_oracle"), o.get("error"))
show(d)
for o in d.get("other_workloads",[]): show(o)
print("skipped", d.get("other_workloads_skipped"), "launches", d.get("gpu_launches"))
PY
