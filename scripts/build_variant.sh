#!/bin/bash
# Development tool: build libffgpu with extra compiler flags into build/<name>.so for same-box A/B runs
# (FFGPU_LIB=$PWD/build/<name>.so python scripts/kbench.py ...).   usage: build_variant.sh <name> [nvcc flags...]
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o build/$name.so firefly_b200/csrc/capi.cu firefly_b200/csrc/kernels.cu firefly_b200/csrc/compiler.cpp firefly_b200/csrc/queue.cpp
ls -la build/$name.so
