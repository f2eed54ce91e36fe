#!/bin/bash
# Development helper (GPU box): run a command once per library variant.  usage: scripts/ab.sh "<cmd>" base v1 v2 ...
cmd=$1; shift
lib=continual-learning-for-dynamic-video-quality-enhancement_b200/nerve_cl_b200/libnervecl.so
cp $lib /tmp/base.so
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/base.so $lib; else cp variants/$v.so $lib; fi
  echo "== $v"; eval "$cmd"
done
cp /tmp/base.so $lib
