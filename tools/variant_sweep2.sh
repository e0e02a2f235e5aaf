#!/bin/bash
cd /root/repo
for v in "$@"; do
  for w in C3 C4; do
  GNNMP_LIB_PATH=/root/repo/gnn_motion_planning_b200/libgnnmp_$v.so python bench.py --workload $w --steps 5 --no-sub-records --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']; print('$v $w', 'step %.2f collision %.3f' % (d['ms_per_step'], p['collision']))"
  done
done
