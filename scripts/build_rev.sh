#!/bin/bash
# builds qmcpack_b200/libqmcb_<name>.so from the csrc/ of a git revision (same flags as build_variant.sh) for same-box A/B runs:
#   scripts/build_rev.sh <name> <rev> [extra nvcc flags]
set -e
NAME=$1; REV=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
mkdir -p $TMP/qmcpack_b200/csrc $TMP/include
for f in $(git -C $ROOT ls-tree --name-only $REV qmcpack_b200/csrc/ include/); do git -C $ROOT show $REV:$f > $TMP/$f; done
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 -ccbin /usr/bin/g++"
for f in spline.cu crowd.cu api.cu vmc_host.cpp dmc_host.cpp; do
  /usr/local/cuda/bin/nvcc $FLAGS "$@" -c $TMP/qmcpack_b200/csrc/$f -o $TMP/${f%.*}.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o $ROOT/qmcpack_b200/libqmcb_$NAME.so $TMP/*.o -lcublas -ccbin /usr/bin/g++ -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
rm -rf $TMP
echo $ROOT/qmcpack_b200/libqmcb_$NAME.so
