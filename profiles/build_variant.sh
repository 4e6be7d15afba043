#!/bin/bash
# build a tuning variant of the library: profiles/build_variant.sh <name> <extra nvcc flags...>
# -> duckdb.mbt_b200/csrc/variants/lib_<name>.so   (select with DMB_LIB_PATH)
set -e
cd "$(dirname "$0")/../duckdb.mbt_b200/csrc"
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -I../../include -I. "$@" \
  -shared -o variants/lib_$name.so *.cu
echo built variants/lib_$name.so
