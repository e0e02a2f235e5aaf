#!/bin/bash
# on the GPU box: time the C2 edge-feature phase with each libgnnmp_*.so variant, in modes tc and tcrd
cd /root/repo
for v in "$@"; do
  for m in tc tcrd; do
    lib=/root/repo/gnn_motion_planning_b200/libgnnmp_$v.so
    if [ "$v" = "base" ]; then lib=/root/repo/gnn_motion_planning_b200/libgnnmp.so; fi
    GNNMP_LIB_PATH=$lib timeout 300 python bench.py --steps 10 --no-sub-records --no-cpu-baseline --ef-mode $m 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']; print('$v $m', 'step %.2f edge_feature %.3f' % (d['ms_per_step'], p['edge_feature']))"
  done
done
