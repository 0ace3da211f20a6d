#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/models.log
for t in tests/test_unet_gpu.py tests/test_pipeline_gpu.py; do
  echo "=== $t" >> gpurun_out/models.log
  timeout 900 python -m pytest "$t" -q -s --no-header -p no:cacheprovider 2>&1 | tail -80 >> gpurun_out/models.log
done
grep -E "rel err|passed|failed|Error|error|assert" gpurun_out/models.log | tail -60
