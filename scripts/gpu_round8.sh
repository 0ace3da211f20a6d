#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_processors_gpu.py -m gpu -x -q --no-header -p no:cacheprovider -k "attention or processor or ip" 2>&1 | tail -5
timeout 300 python scripts/attn_bench.py 2>&1 | grep -E "attn"
