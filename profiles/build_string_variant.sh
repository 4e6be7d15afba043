#!/bin/bash
# a tuning variant that differs only in kernels_string.cu: profiles/build_string_variant.sh <name> <extra nvcc flags...>
# -> duckdb.mbt_b200/csrc/variants/lib_<name>.so (select with DMB_LIB_PATH); the other objects come from csrc/build/
set -e
cd "$(dirname "$0")/../duckdb.mbt_b200/csrc"
name=$1; shift
mkdir -p variants build/var_$name
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -I../../include -I. "$@" \
  -c -o build/var_$name/kernels_string.o kernels_string.cu
objs=$(ls build/*.o | grep -v kernels_string.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_$name.so $objs build/var_$name/kernels_string.o -Xcompiler -pthread
echo built variants/lib_$name.so
