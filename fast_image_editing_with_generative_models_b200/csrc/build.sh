#!/bin/bash
# Builds libfie_b200.so (sm_100a) in-tree. Usage: csrc/build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libfie_b200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRCS="capi.cu canny.cu resize.cu elementwise.cu norm.cu conv_small.cu gemm_conv.cu attention.cu attn_vae.cu jpeg.cu pack.cu"
mkdir -p "$HERE/_obj"
pids=()
for s in $SRCS; do
  ( cd "$HERE" && $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c $s -o _obj/${s%.cu}.o ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
cd "$HERE" && $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $(for s in $SRCS; do echo _obj/${s%.cu}.o; done) -lcudart
echo "built $OUT"
