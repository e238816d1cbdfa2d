#!/bin/bash
# one GPU session: tests, benches, launch list, targeted ncu captures.  Everything lands in gpurun_out/.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
if [ "${RUN_TESTS:-1}" = "1" ]; then
  timeout 1200 python -m pytest tests -q -m gpu --timeout 120 --timeout-method=thread --maxfail=10 --no-header -rf 2>&1 | tail -60 > gpurun_out/tests.log
  echo "== tests exit $?" >> $S; grep -E "passed|failed|error" gpurun_out/tests.log | tail -3 >> $S
fi
timeout 300 python tools/bench_vq.py > gpurun_out/bench_vq.json 2> gpurun_out/bench_vq.err; echo "== bench_vq exit $?" >> $S; cat gpurun_out/bench_vq.json >> $S
for W in ${WORKLOADS:-cfg64}; do
  timeout 900 python bench.py --workload $W --steps ${STEPS:-3} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  echo "== bench $W exit $?" >> $S; cat gpurun_out/bench_$W.json >> $S; tail -5 gpurun_out/bench_$W.err >> $S
done
if [ "${RUN_NCU:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80000 --csv --log-file gpurun_out/launches.csv \
     python bench.py --workload cfg64 --quick --steps 1 --warmup 1 > gpurun_out/ncu_launches.log 2>&1
  echo "== ncu launches exit $?" >> $S
  python tools/summarise_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -40 gpurun_out/launches_summary.txt >> $S
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 30 -c 6 -f -o gpurun_out/prof_gemm \
     python bench.py --workload cfg64 --quick --steps 1 --warmup 0 --segment-length 3 > gpurun_out/ncu_gemm.log 2>&1; echo "== ncu gemm exit $?" >> $S
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_argmin_kernel -c 2 -f -o gpurun_out/prof_vq \
     python tools/bench_vq.py > gpurun_out/ncu_vq.log 2>&1; echo "== ncu vq exit $?" >> $S
fi
cat $S
