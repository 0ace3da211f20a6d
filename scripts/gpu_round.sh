#!/bin/bash
# unit tests (attention/norm kernels changed) -> kernel bench -> bench -> ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -x -q --no-header -p no:cacheprovider 2>&1 | tail -8
timeout 600 python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; grep -E "attn|groupnorm|layernorm" gpurun_out/kernel_bench.log
timeout 1500 python bench.py --steps 2 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-1500
tail -3 gpurun_out/bench.err
bash scripts/gpu_profile.sh 2>&1 | tail -8
