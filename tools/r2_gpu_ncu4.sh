#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/cfg4_min.py 6"
$CMD > gpurun_out/ncu4_plain.log 2>&1 || exit 1
cat gpurun_out/ncu4_plain.log
ncu --set full --clock-control none --import-source on -k regex:yolo_decode_filter_kernel -s 4 -c 1 -o gpurun_out/r2_decode_cfg4 $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:sort_nms_kernel -s 4 -c 1 -o gpurun_out/r2_sortnms_cfg4 $CMD > /dev/null 2>&1
ls -la gpurun_out/*cfg4*.ncu-rep
