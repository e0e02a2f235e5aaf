#!/bin/bash
# build_variant.sh NAME "-DFLAG=..."  ->  gnn_motion_planning_b200/libgnnmp_NAME.so (explorer.cu rebuilt with the flags, other objects reused)
set -e
cd /root/repo/gnn_motion_planning_b200/csrc
mkdir -p build/var_$1
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a $2 -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall --expt-relaxed-constexpr -c explorer.cu -o build/var_$1/explorer.o 2> build/var_$1/log.txt
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libgnnmp_$1.so build/api.o build/maze.o build/knn.o build/var_$1/explorer.o build/reduce.o build/smoother.o build/arm.o -lcudart
echo built $1
