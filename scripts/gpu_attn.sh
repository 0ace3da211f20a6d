#!/bin/bash
mkdir -p gpurun_out
python scripts/debug_attn.py 2>&1 | grep "max err"
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q --no-header -p no:cacheprovider -k "attention" 2>&1 | tail -4
timeout 600 python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; grep -E "attn" gpurun_out/kernel_bench.log
