#!/bin/bash
# first GPU bring-up: run every gpu test file separately (no -x) under hard timeouts, keep all logs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for t in test_vq_argmin test_gemm_conv test_elementwise test_llama test_tokenizer; do
  timeout 600 python -m pytest tests/$t.py -q -m gpu --timeout 120 -x --no-header 2>&1 | tail -40 > gpurun_out/$t.log
  echo "== $t exit $?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/$t.log >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
