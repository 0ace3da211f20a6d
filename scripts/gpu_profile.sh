#!/bin/bash
mkdir -p gpurun_out
# launch list of the 2nd forward (first is warm-up): skip setup kernels by NVTX range filtering
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "unet_forward/" -c 1200 --csv --log-file gpurun_out/launches_unet.csv python scripts/profile_unet.py 1 > gpurun_out/prof1.log 2>&1
tail -2 gpurun_out/prof1.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --nvtx --nvtx-include "unet_forward/" -s 120 -c 4 -o gpurun_out/prof_gemm -f python scripts/profile_unet.py 1 > gpurun_out/prof2.log 2>&1
tail -2 gpurun_out/prof2.log
ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel --nvtx --nvtx-include "unet_forward/" -s 60 -c 2 -o gpurun_out/prof_attn -f python scripts/profile_unet.py 1 > gpurun_out/prof3.log 2>&1
tail -2 gpurun_out/prof3.log
ls -la gpurun_out/*.ncu-rep
