#!/bin/bash
# same-box A/B of decode megakernel builds / run-time paths (wall time of a 237-token rollout, B=64):
# the session-start tree (ab/build/old, made by tools/build_variants.sh) against HEAD's paths.
mkdir -p gpurun_out; OUT=gpurun_out/variants.txt; : > $OUT
run() { echo "== $1" >> $OUT; shift; "$@" >> $OUT 2>> gpurun_out/variants.err; }
export PHASES=${PHASES:-0}
[ -d ab/build/old ] && run "old tree (aceeb50)" env MEGA_ONLY=1 python ab/build/old/tools/mega_profile.py
run "HEAD" env VARIANTS="${VARIANTS:-0:0:0,0:1:0,0:0:1,0:1:1,1:1:1}" python tools/mega_profile.py
for v in ${LIBS}; do
  run "HEAD + $v" env VARIANTS="${VARIANTS_LIB:-0:1:1}" IVGPT_B200_LIB=$PWD/ab/build/lib_$v.so python tools/mega_profile.py
done
cat $OUT
