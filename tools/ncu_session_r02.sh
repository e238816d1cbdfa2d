#!/bin/bash
# Round-2 profiler evidence (one GPU box): launch list of the bench command, `ncu --set full` of the decode megakernel (16-step
# launch), of three representative conv launches and of the prefill's fused-attention kernel.  The raw / source pages are
# exported to CSV on the box (the .ncu-rep files are too big to bring back); summaries are made by tools/ncu_summarise_r02.py.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80000 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --workload cfg64 --quick --steps 2 --warmup 1 > gpurun_out/r02_ncu_launches.log 2>&1
python tools/summarise_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_cfg64.txt 2>&1
rm -f gpurun_out/r02_launches.csv
NEW=17 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega_kernel -c 1 -f -o /tmp/prof_mega \
   python tools/mega_ncu.py > gpurun_out/r02_ncu_mega.log 2>&1
ncu -i /tmp/prof_mega.ncu-rep --page raw --csv > gpurun_out/r02_ncu_decode_mega_raw.csv 2>/dev/null
SHAPES=0,1,2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 6 -c 1 -f -o /tmp/prof_conv0 \
   python tools/conv_profile.py > gpurun_out/r02_ncu_conv.log 2>&1
SHAPES=1 timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 6 -c 1 -f -o /tmp/prof_conv1 python tools/conv_profile.py >> gpurun_out/r02_ncu_conv.log 2>&1
SHAPES=2 timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 6 -c 1 -f -o /tmp/prof_conv2 python tools/conv_profile.py >> gpurun_out/r02_ncu_conv.log 2>&1
for i in 0 1 2; do ncu -i /tmp/prof_conv$i.ncu-rep --page raw --csv > gpurun_out/r02_ncu_conv_shape${i}_raw.csv 2>/dev/null; done
NEW=1 timeout 600 ncu --set full --clock-control none -k regex:flash_attn_kernel -s 12 -c 1 -f -o /tmp/prof_fa python tools/mega_ncu.py > gpurun_out/r02_ncu_fa.log 2>&1
ncu -i /tmp/prof_fa.ncu-rep --page raw --csv > gpurun_out/r02_ncu_flash_attn_raw.csv 2>/dev/null
ls -la gpurun_out/r02_*
