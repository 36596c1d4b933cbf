#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 6 --warmup 3 --no-configs --e2e-steps 2 --depth 1"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench_cfg2.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:yolo_decode_filter_kernel -s 4 -c 1 -o gpurun_out/r2_decode_b256 $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:small_nms_kernel -s 4 -c 1 -o gpurun_out/r2_small_nms_b256 $CMD > /dev/null 2>&1
CMD32="python bench.py --batch 32 --steps 6 --warmup 3 --no-configs --e2e-steps 2 --depth 1"
$CMD32 > gpurun_out/ncu_plain32.log 2>&1 && ncu --set full --clock-control none -k regex:yolo_decode_filter_kernel -s 4 -c 1 -o gpurun_out/r2_decode_b32 $CMD32 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
