#!/bin/bash
# Experiment build of libavzoom (reads AVZ_IBM_TOL, AVZ_FUSED_* from the environment): gpurun_tmp/libavzoom_exp.so
# usage: tools/build_exp.sh [extra nvcc flags, e.g. -DAVZ_COV_NOFFT] ; use with AVZ_LIB=$PWD/gpurun_tmp/libavzoom_exp.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/real-time-audio-visual-zooming_b200/csrc
OUT=${AVZ_EXP_OUT:-$ROOT/gpurun_tmp/libavzoom_exp.so}
TMP=$(mktemp -d)
for f in avz_host avz_generic avz_pointwise avz_opt512 avz_opt1024 avz_mixer avz_precise; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden \
       -I$ROOT/include -I$SRC --expt-relaxed-constexpr -DAVZ_EXPERIMENT "$@" -c $SRC/$f.cu -o $TMP/$f.o &
done
wait
mkdir -p $(dirname $OUT)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $TMP/*.o -cudart static
rm -rf $TMP
ls -la $OUT
