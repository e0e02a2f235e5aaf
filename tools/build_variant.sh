#!/bin/bash
# build_variant.sh NAME "-DFLAG=..." [source.cu]  ->  gnn_motion_planning_b200/libgnnmp_NAME.so (one source rebuilt with the flags, other objects reused)
set -e
SRC=${3:-explorer.cu}
OBJ=${SRC%.cu}
cd /root/repo/gnn_motion_planning_b200/csrc
mkdir -p build/var_$1
EXTRA=""
if [ "$SRC" = "arm.cu" ]; then EXTRA="-fmad=false"; fi
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a $2 $EXTRA -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall --expt-relaxed-constexpr -c $SRC -o build/var_$1/$OBJ.o 2> build/var_$1/log.txt
OBJS=""
for o in api maze knn explorer reduce smoother arm; do
  if [ "$o" = "$OBJ" ]; then OBJS="$OBJS build/var_$1/$o.o"; else OBJS="$OBJS build/$o.o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libgnnmp_$1.so $OBJS -lcudart
echo built $1
