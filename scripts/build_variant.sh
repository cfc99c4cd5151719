#!/bin/bash
# Build a variant of the library with extra -D flags: bash scripts/build_variant.sh variants/lib_x.so -DGG_EARLY_TICKET=0 ...
out=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -shared -o $out \
  gargammel_b200/csrc/gg_api.cu gargammel_b200/csrc/gg_kernels.cu gargammel_b200/csrc/gg_host.cpp -lz
