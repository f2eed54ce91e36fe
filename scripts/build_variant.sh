#!/bin/bash
# Development helper: build libnervecl.so with extra -D flags for ONE source file into variants/<name>.so
# usage: scripts/build_variant.sh <name> <file.cu> -DFOO=1 ...
set -e
name=$1; src=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
csrc=$root/continual-learning-for-dynamic-video-quality-enhancement_b200/csrc
mkdir -p $root/variants $csrc/build
make -C $csrc -j16 > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden \
  -I$root/include --expt-relaxed-constexpr "$@" -c $csrc/$src -o /tmp/variant_$name.o
objs=$(ls $csrc/build/*.o | grep -v "/${src%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/variants/$name.so $objs /tmp/variant_$name.o -lcudart_static -ldl -lrt -lpthread
echo built variants/$name.so
