#!/bin/bash
# a tuning variant that differs in ONE source file: profiles/build_one_variant.sh <file-stem> <name> <extra nvcc flags...>
# -> duckdb.mbt_b200/csrc/variants/lib_<name>.so (select with DMB_LIB_PATH); the other objects come from csrc/build/
set -e
cd "$(dirname "$0")/../duckdb.mbt_b200/csrc"
stem=$1; name=$2; shift; shift
mkdir -p variants build/var_$name
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -I../../include -I. "$@" \
  -c -o build/var_$name/$stem.o $stem.cu
objs=$(ls build/*.o | grep -v "/$stem.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_$name.so $objs build/var_$name/$stem.o -Xcompiler -pthread
echo built variants/lib_$name.so
