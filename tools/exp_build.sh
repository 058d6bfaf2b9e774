#!/bin/bash
# Throw-away variant of the library: one source recompiled with extra -D flags, linked with the stock objects.
#   tools/exp_build.sh <tag> <file.cu> <flags...>   ->  maxpro_b200/libmaxpro_b200_<tag>.so
set -e
cd "$(dirname "$0")/../maxpro_b200/csrc"
TAG=$1; SRC=$2; shift 2
mkdir -p build_exp
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -warn-spills"
$NV "$@" -c $SRC -o build_exp/${SRC%.cu}_$TAG.o
OBJS=$(ls build/*.o | grep -v "/${SRC%.cu}.o")
$NV -shared -o ../libmaxpro_b200_$TAG.so $OBJS build_exp/${SRC%.cu}_$TAG.o
echo built ../libmaxpro_b200_$TAG.so
