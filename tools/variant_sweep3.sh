#!/bin/bash
cd /root/repo
for rep in 1 2; do
for v in "$@"; do
  L=/root/repo/gnn_motion_planning_b200/libgnnmp_$v.so
  if [ "$v" = "base" ]; then L=/root/repo/gnn_motion_planning_b200/libgnnmp.so; fi
  for w in C2 C4; do
  GNNMP_LIB_PATH=$L python bench.py --workload $w --steps 8 --no-sub-records --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']; print('$v $w', 'step %.2f edge_feature %.3f edge_msg %.3f' % (d['ms_per_step'], p['edge_feature'], p['edge_msg']))"
  done
done
done
