#!/bin/bash
# final 1-GPU evidence of a commit: GPU tests, smoke, the reference arm and the default bench line (as the driver runs them),
# the c3 launch list and the shard timings.   bash tools/gpu_final.sh [tag]
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt; tail -3 $OUT/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "exit $?" >> $OUT/smoke.txt; tail -2 $OUT/smoke.txt
bash tools/gpu_bench_only.sh $TAG
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench_c3.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-others --no-verify > $OUT/launches_bench_c3.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_band500.csv python tools/small_band_once.py > $OUT/band500.log 2>&1
timeout 300 python tools/small_band.py > $OUT/small_band.txt 2>&1; head -14 $OUT/small_band.txt
