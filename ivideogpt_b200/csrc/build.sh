#!/bin/bash
# Builds libivgpt_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --use_fast_math=false"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p build
pids=()
for f in capi vq_argmin gemm_tc elementwise llama_ops decode_mega llama_train flash_attn tok_train; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ gemm_params.cuh -nt build/$f.o ] || [ decode_mega.cuh -nt build/$f.o ] || [ ../../include/ivgpt_b200.h -nt build/$f.o ]; then
    $NVCC $FLAGS -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o ../libivgpt_b200.so build/capi.o build/vq_argmin.o build/gemm_tc.o build/elementwise.o build/llama_ops.o build/decode_mega.o build/llama_train.o build/flash_attn.o build/tok_train.o -lcudart
echo "built $(cd .. && pwd)/libivgpt_b200.so"
