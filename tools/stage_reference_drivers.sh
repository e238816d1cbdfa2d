#!/bin/bash
# Stages UNMODIFIED copies of the reference's driver scripts (and one sample episode) under baseline/_ref/ so that
# tests/test_dropin_drivers.py can execute them on the GPU box, where /root/reference does not exist.  baseline/_ref/ is
# git-ignored (never part of the history, never product code) but travels with gpurun snapshots -- the same arrangement the
# bench contract uses for the reference arm.  Run in the build container only.
set -e
cd "$(dirname "$0")/.."
REF=${REF:-/root/reference}
mkdir -p baseline/_ref/inference/samples
cp "$REF/inference/predict.py" "$REF/inference/utils.py" baseline/_ref/inference/
cp "$REF/inference/samples/bair_sample.npz" baseline/_ref/inference/samples/
cp "$REF/train_gpt.py" "$REF/train_tokenizer.py" baseline/_ref/
( cd "$REF" && sha256sum inference/predict.py inference/utils.py train_gpt.py train_tokenizer.py ) > baseline/_ref/SHA256SUMS
echo "staged: $(find baseline/_ref -type f | wc -l) files"
