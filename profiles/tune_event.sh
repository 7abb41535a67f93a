#!/bin/bash
# sweep of the cooperative event kernel's launch shape (blocks per SM) and of the block-local finishing threshold
for bps in 1 2; do for loc in 1024 4096 8192 32768; do
  MCAC_B200_COOP_BPS=$bps MCAC_B200_SORT_LOCAL=$loc python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('bps=$bps local=$loc value=%.0f event_us=%.0f commit_us=%.0f cells_us=%.0f' % (d['value'], r['avg_launch_us'], r['avg_commit_us'], 0.0))"
done; done
