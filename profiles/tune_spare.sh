#!/bin/bash
# SMs left free by the event kernel for the overlapped cell rebuild
for n in 0 4 8 16 32; do
  MCAC_B200_EVENT_SPARE_SMS=$n python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-kernel-table 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('spare_sms=$n value=%.0f event_us=%.1f search_us=%.1f commit_us=%.1f' % (d['value'], r['avg_launch_us'], r['avg_search_us'], r['avg_commit_us']))"
done
