#!/bin/bash
mkdir -p gpurun_out
for w in main fused; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --nvtx --nvtx-include "unet_forward/" --csv --log-file gpurun_out/launches_$w.csv python scripts/profile_fused.py $w > gpurun_out/prof_$w.log 2>&1
tail -1 gpurun_out/prof_$w.log
done
