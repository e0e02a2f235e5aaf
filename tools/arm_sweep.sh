#!/bin/bash
# on the GPU box: rebuild arm.cu with different occupancy targets and time the C4 collision phase
cd /root/repo
for mb in 4 5 6 8; do
  touch gnn_motion_planning_b200/csrc/arm.cu
  make -C gnn_motion_planning_b200/csrc -s EXTRA=-DGMP_ARM_MINB=$mb > /dev/null 2>&1
  grep -A3 "arm_edge_graph_fast_kernel" gnn_motion_planning_b200/csrc/build/arm.ptxas.log | tail -2 | tr '\n' ' '
  python bench.py --workload C4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(' MINB=$mb collision ms', d['phases_ms_per_step']['collision'])"
done
