#!/bin/bash
# A/B inside one box: CHB_INLINE_MIN (uncertified pixels per tile from which the tile is finished inside the streaming kernel)
for v in 12 4 8 20 33; do
  CHB_INLINE_MIN=$v timeout 300 python bench.py --workload c3-outlier-abs-extreme --no-cpu --no-e2e --no-others --no-verify 2>/dev/null | python -c "
import json,sys
o=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=o['roofline']
print('inline_min $v c3 ms %.4f call %.4f main %.4f tiers %.4f' % (o['ms_per_step'], r['avg_launch_ms'], r['dominant_kernel']['avg_launch_ms'], r['tier_kernels_ms']))"
done
for v in 12 33; do
  CHB_INLINE_MIN=$v timeout 300 python bench.py --workload a4-gauss-noise --no-cpu --no-e2e --no-others --no-verify 2>/dev/null | python -c "
import json,sys
o=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=o['roofline']
print('inline_min $v a4 ms %.4f call %.4f main %.4f tiers %.4f' % (o['ms_per_step'], r['avg_launch_ms'], r['dominant_kernel']['avg_launch_ms'], r['tier_kernels_ms']))"
done
