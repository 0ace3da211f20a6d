#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -12
timeout 600 python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; grep -E "tall|single|conv3x3" gpurun_out/kernel_bench.log | cut -c1-150
timeout 1500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -1 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','unet_step_ms','whole_path_tflops','gpu_launches')}, d['roofline']['achieved'], d['e2e']['value'])
PY
