#!/bin/sh
# debug build of the library with device-side printf (never shipped)
cd "$(dirname "$0")/../fqzcomp5_b200" && nvcc -gencode arch=compute_100a,code=sm_100a -O1 -lineinfo -std=c++17 -fmad=false -DB200_DEBUG \
  -Xcompiler -fPIC,-fvisibility=hidden --cudart=static -shared -o libb200rans.so csrc/kernels.cu csrc/stripe.cu csrc/api.cu -lpthread -ldl -lrt
