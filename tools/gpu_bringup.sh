#!/bin/bash
# GPU bring-up: every gpu test file separately under hard timeouts, all logs kept in gpurun_out/
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python tools/diag_gemm.py > gpurun_out/diag_gemm.log 2>&1; echo "== diag exit $?" >> gpurun_out/summary.txt
tail -30 gpurun_out/diag_gemm.log >> gpurun_out/summary.txt
for t in ${TESTS:-test_vq_argmin test_gemm_conv test_elementwise test_llama test_tokenizer}; do
  timeout 900 python -m pytest tests/$t.py -q -m gpu --timeout 90 --timeout-method=thread --maxfail=8 --no-header -rf 2>&1 | tail -60 > gpurun_out/$t.log
  echo "== $t exit $?" >> gpurun_out/summary.txt
  grep -E "passed|failed|error" gpurun_out/$t.log | tail -3 >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
