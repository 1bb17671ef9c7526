#!/bin/bash
# build an A/B variant of the library: scratch/build_variant.sh NAME "-DFLAG ..."  -> scratch/libfest3d_gpu_NAME.so (select with F3D_LIB)
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../fest-3d_b200/csrc"
tmp=/tmp/f3d_variant_$name; mkdir -p $tmp
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags"
for f in api sweep sweep3 sweep3_rare bc grad util walldist checkpoint; do $NV -c $f.cu -o $tmp/$f.o & done
$NV -fmad=false -c geometry.cu -o $tmp/geometry.o &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/libfest3d_gpu_$name.so $tmp/*.o -lcudart -ldl -lpthread
echo built scratch/libfest3d_gpu_$name.so
