#!/bin/bash
# 2-GPU run: multi-GPU parity test + the bench under torchrun (strong scaling on interleaved row blocks, NCCL gather, verification)
TAG=${1:-r2y}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L > $OUT/gpus.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "multi_gpu or interleaved" > $OUT/pytest_multi.txt 2>&1; tail -3 $OUT/pytest_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-others > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
tail -5 $OUT/bench_2gpu.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_2gpu.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "scaling", d["scaling"], "run", d["run"]["sharding"])
print("e2e", json.dumps(d.get("e2e"))[:500])
print("verified", d.get("verified"))
print("weak", d.get("weak_scaling"))
PY
CHB_BENCH_CONTIGUOUS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-others --no-e2e --no-verify > $OUT/bench_2gpu_contiguous.json 2> $OUT/bench_2gpu_contiguous.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_2gpu_contiguous.json").read().strip().splitlines()[-1])
print("contiguous bands: value", d["value"], "ms", d["ms_per_step"], d["run"]["sharding"])
PY
