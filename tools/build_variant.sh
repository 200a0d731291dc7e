#!/bin/bash
# Experiments: build keyword_spotting_b200/variants/libkws_<name>.so from the cached objects, with ONE source
# recompiled with extra flags.  usage: tools/build_variant.sh <name> <source.cu> <nvcc flags...>
# Run with KWS_B200_LIB=keyword_spotting_b200/variants/libkws_<name>.so to select it.
set -e
cd "$(dirname "$0")/../keyword_spotting_b200"
name=$1; src=$2; shift 2
mkdir -p variants /tmp/kws_variants
obj=/tmp/kws_variants/${name}_$(basename $src .cu).o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c csrc/$src -o $obj
others=$(ls build/*.o | grep -v "/$(basename $src .cu).o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o variants/libkws_${name}.so $obj $others
echo variants/libkws_${name}.so
