#!/bin/bash
# A/B inside one box: video_kernel launch bounds
export PYTHONUNBUFFERED=1
one() {  # lib workload
  CHB_LIB=$1 timeout 300 python bench.py --workload $2 --no-cpu --no-e2e --no-others --no-verify 2>/dev/null | python -c "
import json,sys
o=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=o.get('roofline') or {}
print('$1 $2 ms %.3f frac %.3f tiers %.3f' % (o['ms_per_step'], r.get('frac',0), r.get('tier_kernels_ms') or 0))"
}
B=$PWD/chrono_photo_b200/libchrono_b200.so
V=$PWD/chrono_photo_b200/_variants
for rep in 1 2; do
  one $B c5-video; one $V/vminb3.so c5-video
done
