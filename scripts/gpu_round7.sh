#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -3
for shp in "4 4096 10" "4 1024 20"; do
  tag=$(echo $shp | tr ' ' '_')
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_tc --launch-skip 2 --launch-count 1 \
     -o gpurun_out/r01_attn_$tag -f python scripts/attn_one.py $shp > gpurun_out/ncu_attn_$tag.log 2>&1
  tail -2 gpurun_out/ncu_attn_$tag.log
done
timeout 600 python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; grep -E "conv3x3_res(128|64)_(bn0|tall)|conv3x3_res128_cat_(bn0|tall)" gpurun_out/kernel_bench.log | cut -c1-150
