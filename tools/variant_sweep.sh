#!/bin/bash
# on the GPU box: time the C2 phases with each libgnnmp_*.so variant given on the command line
cd /root/repo
for v in "$@"; do
  GNNMP_LIB_PATH=/root/repo/gnn_motion_planning_b200/libgnnmp_$v.so python bench.py --steps 10 --no-sub-records --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']; print('$v', 'step %.2f edge_feature %.3f edge_msg %.3f policy %.3f' % (d['ms_per_step'], p['edge_feature'], p['edge_msg'], p['policy']))"
done
