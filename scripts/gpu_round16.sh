#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 300 python scripts/attn_bench.py 2>&1 | grep -E "attn" > gpurun_out/attn_bench.log; cat gpurun_out/attn_bench.log
timeout 1500 python bench.py --steps 2 --warmup 3 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -1 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','unet_step_ms','whole_path_tflops','gpu_launches')}, d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'], d['effective'], d['vae_decode']['ms_per_stage'], d['cpu_baseline']['value'], d['clocks'])
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-400
