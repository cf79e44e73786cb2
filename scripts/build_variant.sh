#!/bin/bash
# builds a tuning / debugging variant of libqmcb.so with extra nvcc flags: scripts/build_variant.sh <name> <flags...>
# -> qmcpack_b200/libqmcb_<name>.so (select with QMCB_LIB=...); the default library is left alone
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/qmcpack_b200/csrc
OUT=$ROOT/qmcpack_b200/libqmcb_$NAME.so
TMP=$(mktemp -d)
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 -ccbin /usr/bin/g++"
for f in spline.cu crowd.cu api.cu vmc_host.cpp dmc_host.cpp; do
  /usr/local/cuda/bin/nvcc $FLAGS "$@" -c $CSRC/$f -o $TMP/${f%.*}.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o $OUT $TMP/*.o -lcublas -ccbin /usr/bin/g++ -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
rm -rf $TMP
echo $OUT
