#!/bin/bash
# Round-1 evidence on one B200: (1) launch list of one bench step, (2) ncu --set full of the dominant kernel.
# Each ncu run only after the same command exited 0 without ncu.
CMD="python bench.py --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/final_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/final_ncu1.log 2>&1
$CMD > gpurun_out/final_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kg_insert_segs_kernel -s 2 -c 2 -o gpurun_out/r01_insert_segs $CMD > gpurun_out/final_ncu2.log 2>&1
CMD2="python bench.py --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --partitions 1"
$CMD2 > gpurun_out/final_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kg_count_kernel -s 2 -c 2 -o gpurun_out/r01_count_direct $CMD2 > gpurun_out/final_ncu3.log 2>&1
tail -c 150 gpurun_out/final_ncu1.log gpurun_out/final_ncu2.log gpurun_out/final_ncu3.log 2>/dev/null | tr '\n' ' '
