#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider -s 2>&1 | grep -E "dedup vs|passed|failed|Error|error|assert" | tail -8
timeout 1500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -2 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','unet_step_ms','whole_path_tflops','gpu_launches')}, d['roofline']['achieved'], d['e2e']['value'], d.get('effective'), d.get('vae_decode'))
PY
