#!/bin/bash
# sweep of K1's group width (lanes per query) and occupancy target on the N=1e6 bench state: one launch of 349 000 searches
for cfg in "0 4" "4 4" "4 6" "4 8" "8 4" "8 6" "8 8" "16 4" "32 4" "-1 6"; do
  set -- $cfg
  MCAC_B200_SEARCH_GROUP=$1 MCAC_B200_SEARCH_MB=$2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline_k1_sweep"]; q=d["roofline_k1_inloop"]; print('group=$1 minblocks=$2 sweep_ms=%.3f searches/s=%.3g GB/s=%.0f frac=%.3f | in-loop search_us=%.1f value=%.0f' % (r['kernel_ms'], r['searches_per_sec'], r['achieved'], r['frac'], q['avg_launch_us'], d['value']))"
done
