#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vae_gpu.py -m gpu -x -q --no-header -p no:cacheprovider -s 2>&1 | tail -15
timeout 600 python scripts/vae_bench.py 2>&1 | tail -3
