#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc --launch-skip 2 --launch-count 1 -o gpurun_out/r01_gemm_tall_out1280 -f python scripts/gemm_one.py 4096 1280 1280 > gpurun_out/ncu_g3.log 2>&1; tail -1 gpurun_out/ncu_g3.log
