#!/bin/bash
OUT=gpurun_out/r2h
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt
tail -4 $OUT/pytest_gpu.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err ) 2>&1 | grep real
grep -i "fail\|error\|Traceback" -A5 $OUT/bench_2gpu.err | head -30
python - <<PY
import json
d=json.loads(open("$OUT/bench_2gpu.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("2gpu", d["scaling"], "value %.4g"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.3f"%r["frac"], "tier_ms", r.get("tier_kernels_ms"), "e2e", d.get("e2e"), "verified", d.get("verified"), "weak", d.get("weak_scaling"))
PY
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/bench_ref_2gpu.json 2> $OUT/bench_ref_2gpu.err ) 2>&1 | grep real
cut -c1-300 $OUT/bench_ref_2gpu.json
timeout 300 python bench.py --workload c3-outlier-abs-extreme --no-cpu --no-e2e --no-verify --steps 20 --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1gpu ms %.4f'%d['ms_per_step'], 'launch %.4f'%d['roofline']['avg_launch_ms'])"
timeout 300 python bench.py --workload c1-minimal --no-cpu --no-e2e --steps 20 --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c1 ms %.4f'%d['ms_per_step'], 'launch %.4f'%d['roofline']['avg_launch_ms'], d['verified'])"
