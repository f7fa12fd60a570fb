#!/bin/bash
# A/B inside one box: programmatic dependent launches of the tier kernels on / off, c3 and c1 and c4
mkdir -p gpurun_out/pdl
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
for wl in c3-outlier-abs-extreme c1-minimal c4-outlier-rel-forward; do for pdl in 1 0 1 0; do
CHB_PDL=$pdl python bench.py --workload $wl --no-e2e --no-cpu --steps 20 --warmup 5 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$wl pdl=$pdl', round(d['ms_per_step'],4), round(r['avg_launch_ms'],4), round(r['dominant_kernel']['avg_launch_ms'],4))"
done; done
for wl in a1-iid-uniform a4-gauss-noise; do python bench.py --workload $wl --no-cpu --e2e-steps 1 --steps 10 --warmup 3 > gpurun_out/pdl/bench_$wl.json 2>gpurun_out/pdl/bench_$wl.err; cut -c1-250 gpurun_out/pdl/bench_$wl.json; done
