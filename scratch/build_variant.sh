#!/bin/bash
# build an alternative library from the same sources with extra compile flags: scratch/build_variant.sh NAME "-DFLAG ..."  ->  fest-3d_b200/libfest3d_gpu_NAME.so
# (selected at run time with F3D_LIB=...; A/B measurements inside one gpurun call)
set -e
name=$1; flags=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
cp -r $root/fest-3d_b200/csrc $tmp/csrc; mkdir -p $tmp/include; cp $root/include/fest3d_gpu.h $tmp/include/
sed -i 's#../../include/fest3d_gpu.h#../include/fest3d_gpu.h#' $tmp/csrc/Makefile $tmp/csrc/*.hpp $tmp/csrc/*.cuh $tmp/csrc/*.cu 2>/dev/null || true
rm -f $tmp/csrc/*.o
make -C $tmp/csrc -j8 EXTRA="$flags" OUT=$root/fest-3d_b200/libfest3d_gpu_$name.so > $tmp/build.log 2>&1 || (tail -20 $tmp/build.log; exit 1)
grep -A3 "k_fusedILi7ELi1ELi2ELb1ELb0" $tmp/csrc/fused.ptxas.log | grep -E "spill|Used" || true
rm -rf $tmp
