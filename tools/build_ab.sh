#!/bin/bash
# usage: tools/build_ab.sh TU NAME "-DMACRO=..." : compiles gguf_b200/csrc/TU.cu with extra defines into build_ab/NAME.o and
# links gguf_b200/libggq_ab_NAME.so from it plus the shipped objects (select at run time with GGQ_SO=...).
set -e
tu=$1; name=$2; defs=$3
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/build_ab
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false --ftz=false --prec-div=true --prec-sqrt=true -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v"
nvcc $F $defs -c $root/gguf_b200/csrc/$tu.cu -o $root/build_ab/$name.o 2> $root/build_ab/$name.log
objs=""
for o in api dequant quant_legacy quant_k rearrange convert host_copy; do
  if [ $o = $tu ]; then objs="$objs $root/build_ab/$name.o"; else objs="$objs $root/gguf_b200/csrc/build/$o.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/gguf_b200/libggq_ab_$name.so $objs -cudart static -Xlinker --exclude-libs=ALL
