#!/bin/bash
# usage: scripts/build_variant.sh <name> -DMACRO=val ...   -> mongeampere_b200/variants/libma_b200_<name>.so (git-ignored)
name=$1; shift
mkdir -p mongeampere_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-extended-lambda --expt-relaxed-constexpr \
  -Xcompiler -fPIC -shared "$@" mongeampere_b200/csrc/ma_b200.cu -o mongeampere_b200/variants/libma_b200_$name.so
