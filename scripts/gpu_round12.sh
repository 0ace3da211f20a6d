#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_config1_gpu.py tests/test_eager_baseline_gpu.py -m gpu -x -q --no-header -p no:cacheprovider -s 2>&1 | grep -v "^Number" | tail -12 | cut -c1-1200
