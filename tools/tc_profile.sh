#!/bin/bash
# on the GPU box: rebuild explorer.cu with the clock64 stage counters and print them for one C2 forward
cd /root/repo
touch gnn_motion_planning_b200/csrc/explorer.cu
make -C gnn_motion_planning_b200/csrc -s EXTRA=-DGMP_TC_PROFILE > /dev/null 2>&1
python bench.py --steps 1 --warmup 3 --no-sub-records --no-cpu-baseline 2>&1 | grep -E "profile|issuer" | tail -14
