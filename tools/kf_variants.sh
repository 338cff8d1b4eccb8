#!/bin/bash
# Build experiment variants of the K-F translation unit (extra -D flags) into lib/exp/libb200bo_<tag>.so.
# usage: tools/kf_variants.sh tag "-DKF2_UNROLL=1" [tag2 "flags2" ...]; run with B200BO_LIB_PATH=<that file>
set -e
PKG=$(dirname "$0")/../jmd-diversity-in-bayesian-optimization_b200
mkdir -p $PKG/lib/exp $PKG/build/exp
while [ $# -ge 2 ]; do
  tag=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=true -Xcompiler -fPIC -I $PKG/../include $flags \
       -c $PKG/csrc/bo_persist.cu -o $PKG/build/exp/bo_persist_$tag.o
  objs=$(ls $PKG/build/*.o | grep -v bo_persist.o)
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart shared -Xlinker -rpath,/usr/local/cuda/lib64 \
       -o $PKG/lib/exp/libb200bo_$tag.so $objs $PKG/build/exp/bo_persist_$tag.o
  echo built $tag
done
