#!/bin/bash
# build_variant.sh <out.so> <extra nvcc flags...>: alternative build of the library for tuning experiments
# (selected at run time with CICLOPE_B200_LIB=<out.so>)
set -e
OUT=$1; shift
cd "$(dirname "$0")/../.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude "$@" \
  -shared -o "$OUT" ciclope_b200/csrc/*.cu -lcudart
