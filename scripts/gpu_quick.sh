#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 600 python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; grep -E "attn|out1280|ff2_1280|out640|res32_bn0|res64_bn0" gpurun_out/kernel_bench.log
timeout 1500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-200
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','unet_step_ms','whole_path_tflops')}, d['roofline']['achieved'], d['e2e']['value'])
PY
