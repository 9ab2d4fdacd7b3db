#!/bin/bash
# Build a variant of the library for kernel experiments: scripts/build_variant.sh <name> <extra nvcc flags...>
# -> scripts/_variants/libr4r_<name>.so  (load with R4R_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../reviews4rec_b200/csrc"
out=../../scripts/_variants; mkdir -p $out/obj_$name
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in api gather conv_simt conv_tc wgrad head head_fused adam shard docplan docs dgrad; do
  if [ -f $f.cu ]; then nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC "$@" -c $f.cu -o $out/obj_$name/$f.o & fi
done
wait
nvcc $ARCH -shared -o $out/libr4r_$name.so $out/obj_$name/*.o -lcudart
echo built $out/libr4r_$name.so
