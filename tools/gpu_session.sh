#!/bin/bash
# one GPU session (HEAD validation): tests, headline bench, megakernel phase breakdown, launch list, ncu --set full of the
# dominant kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out; S=gpurun_out/summary.txt; rm -f $S
t0=$(date +%s)
if [ "${RUN_TESTS:-1}" = "1" ]; then
  timeout 900 python -m pytest tests -q -m gpu --timeout 180 --timeout-method=thread --maxfail=10 --no-header -rf 2>&1 | tail -60 > gpurun_out/tests.log
  echo "== tests exit $? ($(( $(date +%s) - t0 )) s)" >> $S; grep -E "passed|failed|error" gpurun_out/tests.log | tail -3 >> $S
fi
for W in ${WORKLOADS:-cfg64}; do
  timeout 600 python bench.py --workload $W --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  echo "== bench $W exit $? ($(( $(date +%s) - t0 )) s)" >> $S; cat gpurun_out/bench_$W.json >> $S; tail -5 gpurun_out/bench_$W.err >> $S
done
for W in ${QUICK_WORKLOADS}; do
  timeout 600 python bench.py --workload $W --steps 2 --warmup 3 --quick > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  echo "== bench(quick) $W exit $? ($(( $(date +%s) - t0 )) s)" >> $S; cat gpurun_out/bench_$W.json >> $S; tail -5 gpurun_out/bench_$W.err >> $S
done
if [ "${RUN_MEGA_PROFILE:-1}" = "1" ]; then
  timeout 600 python tools/mega_profile.py > gpurun_out/mega_profile.json 2> gpurun_out/mega_profile.err
  echo "== mega_profile exit $? ($(( $(date +%s) - t0 )) s)" >> $S; cat gpurun_out/mega_profile.json >> $S; tail -3 gpurun_out/mega_profile.err >> $S
fi
if [ "${RUN_NCU:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80000 --csv --log-file gpurun_out/launches.csv \
     python bench.py --workload cfg64 --quick --steps 1 --warmup 1 > gpurun_out/ncu_launches.log 2>&1
  echo "== ncu launches exit $? ($(( $(date +%s) - t0 )) s)" >> $S
  python tools/summarise_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -30 gpurun_out/launches_summary.txt >> $S
  NEW=17 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega_kernel -c 1 -f -o gpurun_out/prof_mega \
     python tools/mega_ncu.py > gpurun_out/ncu_mega.log 2>&1; echo "== ncu mega exit $? ($(( $(date +%s) - t0 )) s)" >> $S
fi
cat $S
