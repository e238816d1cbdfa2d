#!/bin/bash
# Builds compile-time variants of libivgpt_b200.so (only decode_mega.cu differs) under ab/build/ for a same-box A/B run:
#   tools/build_variants.sh && gpurun -- 'bash tools/run_variants.sh'
set -e
cd "$(dirname "$0")/.."
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
bash ivideogpt_b200/csrc/build.sh > /dev/null
mkdir -p ab/build
B=ivideogpt_b200/csrc/build
declare -A V=( [plainwait]="-DIVG_MEGA_PLAIN_WAIT" [noslots]="-DIVG_MEGA_NO_SLOTS" [nomarks]="-DIVG_MEGA_NO_GEMM_MARKS"
               [alloff]="-DIVG_MEGA_PLAIN_WAIT -DIVG_MEGA_NO_SLOTS -DIVG_MEGA_NO_GEMM_MARKS" )
for name in "${!V[@]}"; do
  ( $NVCC $FLAGS ${V[$name]} -c ivideogpt_b200/csrc/decode_mega.cu -o ab/build/dm_$name.o &&
    $NVCC -shared -o ab/build/lib_$name.so $B/capi.o $B/vq_argmin.o $B/gemm_tc.o $B/elementwise.o $B/llama_ops.o ab/build/dm_$name.o $B/llama_train.o -lcudart ) &
done
wait
# the session-start tree (commit aceeb50) as a full checkout with its own library
if [ ! -d ab/build/old ]; then
  mkdir -p ab/build/old && git archive aceeb50 | tar -x -C ab/build/old
fi
bash ab/build/old/ivideogpt_b200/csrc/build.sh > /dev/null
ls -la ab/build/*.so ab/build/old/ivideogpt_b200/*.so
