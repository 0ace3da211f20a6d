#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -12
timeout 600 python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; grep -E "pair|groupnorm" gpurun_out/kernel_bench.log
for t in 2 1; do echo "OMG_ATTN_TILES=$t"; OMG_ATTN_TILES=$t timeout 300 python scripts/attn_bench.py; done
echo default; timeout 300 python scripts/attn_bench.py | grep -E "cross|ip"
timeout 1500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -1 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','unet_step_ms','whole_path_tflops','gpu_launches')}, d['roofline']['achieved'], d['e2e']['value'])
PY
