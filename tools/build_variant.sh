#!/bin/bash
# usage: tools/build_variant.sh <name> <nvcc -D flags...>
# Builds build/variants/<name>.so: csrc/ep_binning_tiled.cu recompiled with the flags, the other objects taken from build/
# (run `python -m eventpretrain_b200.build --force` first).  The variants travel to the GPU box with the tree.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --extended-lambda -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --fmad=true"
nvcc $F "$@" -c eventpretrain_b200/csrc/ep_binning_tiled.cu -o build/variants/$name.o
objs=$(ls eventpretrain_b200/build/*.o | grep -v ep_binning_tiled.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/$name.so build/variants/$name.o $objs -Xcompiler -fPIC -lcudart
rm build/variants/$name.o
echo built build/variants/$name.so
