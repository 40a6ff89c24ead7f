#!/bin/bash
# Builds libfie_b200.so (sm_100a) in-tree. Usage: csrc/build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${FIE_OUT:-$HERE/../libfie_b200.so}"     # FIE_OUT: experiment builds (load with FIE_LIB=...)
OBJ="${FIE_OBJ:-_obj}"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRCS="capi.cu canny.cu resize.cu elementwise.cu norm.cu conv_small.cu gemm_conv.cu attention.cu attn_vae.cu jpeg.cu pack.cu metrics.cu"
mkdir -p "$HERE/$OBJ"
pids=()
for s in $SRCS; do
  ( cd "$HERE" && $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c $s -o $OBJ/${s%.cu}.o ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
cd "$HERE" && $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $(for s in $SRCS; do echo $OBJ/${s%.cu}.o; done) -lcudart
echo "built $OUT"
