#!/bin/bash
# Builds compile-time variants of libivgpt_b200.so (only decode_mega.cu differs) under ab/build/ for a same-box A/B run:
#   VARIANT_DEFS="name1=-DFLAG1 name2=-DFLAG2" tools/build_variants.sh && gpurun -- 'LIBS="name1 name2" bash tools/run_variants.sh'
# ab/build/old = the tree of commit $OLD (default: the round-1 session-2 start) with its own library, as the fixed reference.
set -e
cd "$(dirname "$0")/.."
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
bash ivideogpt_b200/csrc/build.sh > /dev/null
mkdir -p ab/build
B=ivideogpt_b200/csrc/build
for def in ${VARIANT_DEFS}; do
  name=${def%%=*}; flags=${def#*=}
  ( $NVCC $FLAGS ${flags//,/ } -c ivideogpt_b200/csrc/decode_mega.cu -o ab/build/dm_$name.o &&
    $NVCC -shared -o ab/build/lib_$name.so $B/capi.o $B/vq_argmin.o $B/gemm_tc.o $B/elementwise.o $B/llama_ops.o ab/build/dm_$name.o $B/llama_train.o $B/flash_attn.o $B/tok_train.o -lcudart ) &
done
wait
OLD=${OLD:-aceeb50}
if [ ! -d ab/build/old ]; then
  mkdir -p ab/build/old && git archive $OLD | tar -x -C ab/build/old
  bash ab/build/old/ivideogpt_b200/csrc/build.sh > /dev/null
fi
ls -la ab/build/*.so ab/build/old/ivideogpt_b200/*.so 2>/dev/null
