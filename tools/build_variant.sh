#!/bin/bash
# usage: tools/build_variant.sh <out.so> [extra nvcc flags...]   -- dev tool: build the library with extra -D flags
out=$1; shift
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared -I include "$@" -o "$out" alpha-zero-general-inflexion_b200/csrc/*.cu
