#!/bin/bash
# 8-GPU box: the bench under torchrun at N = 8 (with e2e + gather + verification) and N = 4 (kernel-only)
TAG=${1:-scale8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L > $OUT/gpus.txt
run() {  # n extra-args
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29520 + $1)) bench.py --gpus $1 --steps 10 --warmup 3 --no-others ${@:2} > $OUT/bench_$1gpu.json 2> $OUT/bench_$1gpu.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$1gpu.json").read().strip().splitlines()[-1])
    print("N=$1 value %.4g ms %.4f"%(d["value"], d["ms_per_step"]), d["scaling"], "| weak", (d.get("weak_scaling") or {}).get("ms_per_step"), "| e2e", (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("gather_ms"), "| verified", (d.get("verified") or {}).get("ok"), (d.get("verified") or {}).get("gathered_vs_single_gpu_pixels_differing"))
except Exception as e:
    print("N=$1 failed", e); print(open("$OUT/bench_$1gpu.err").read()[-1500:])
PY
}
run 8
run 4 --no-e2e --no-verify
