#!/bin/bash
# build a variant of the library with extra -D flags for clair_pairs.cu: scratch/build_variant.sh <name> "<flags>"
set -e
name=$1; flags=$2
cd /root/repo/clair_torch_b200/csrc
mkdir -p ../../build/var_$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $flags -c clair_pairs.cu -o ../../build/var_$name/clair_pairs.o
objs=$(ls ../../build/csrc/*.o | grep -v clair_pairs.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/bin/libclair_$name.so $objs ../../build/var_$name/clair_pairs.o
echo built scratch/bin/libclair_$name.so
