#!/bin/bash
# integration/check.sh -- compile checks of the reference-side adapters (no GPU, nothing is run):
#   1. DelayedUpdateB200.h against the reference's update-engine concept (needs /root/reference; g++ -fsyntax-only with
#      the stub config.h of oracle/stub, the recipe of SURVEY.md App. B)
#   2. smoke_main.cpp (a C++ caller of the C ABI, no Python) compiles and links against qmcpack_b200/libqmcb.so
# usage: integration/check.sh [reference_root]      exit code 0 = all checks that could run passed
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(dirname "$HERE")
REF=${1:-/root/reference}
CXX=${CXX:-g++}
if [ -d "$REF/src" ]; then
  R=$REF/src
  $CXX -std=c++17 -fsyntax-only -fopenmp -Wall -Wno-unused-variable -Wno-unused-parameter -Wno-sign-compare -Wno-unknown-pragmas \
    -I"$ROOT/oracle/stub" -I"$ROOT/include" -I"$HERE" -I$R -I$R/Platforms -I$R/Containers -I$R/Utilities -I$R/io \
    -I$R/Particle -I$R/QMCWaveFunctions -I"$REF/external_codes/boost_multi/multi/include" \
    "$HERE/check_engine_concept.cpp"
  echo "engine concept check: ok (DelayedUpdateB200 == DelayedUpdateBatched interface for double, float, complex<double>)"
  $CXX -std=c++17 -fsyntax-only -fopenmp -Wall -Wno-unused-variable -Wno-unused-parameter -Wno-sign-compare -Wno-unknown-pragmas \
    -I"$ROOT/oracle/stub" -I"$ROOT/include" -I"$HERE" -I$R -I$R/Platforms -I$R/Containers -I$R/Utilities \
    "$HERE/check_spline_adapter.cpp"
  echo "spline adapter check: ok (SplineB200Core against multi_UBspline_3d_{s,d} and the reference's containers)"
else
  echo "engine concept check: skipped ($REF/src not found)"
fi
if [ -f "$ROOT/qmcpack_b200/libqmcb.so" ]; then
  mkdir -p "$HERE/_build"
  $CXX -std=c++17 -O2 -Wall -I"$ROOT/include" "$HERE/smoke_main.cpp" -o "$HERE/_build/qmcb_smoke" \
    -L"$ROOT/qmcpack_b200" -lqmcb -Wl,-rpath,"$ROOT/qmcpack_b200" -Wl,-rpath,'$ORIGIN/../../qmcpack_b200'
  echo "C++ smoke harness: built integration/_build/qmcb_smoke"
else
  echo "C++ smoke harness: skipped (qmcpack_b200/libqmcb.so not built)"
fi
