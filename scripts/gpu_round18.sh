#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 600 python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; grep -E "\"(out1280|ff2_1280|ff1_1280|qkv_1280|ff2_640|out640)|tall_bn160|pair_|conv3x3_res(128|64|32)_bn0|geglu" gpurun_out/kernel_bench.log | cut -c1-150
timeout 1500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -1 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','unet_step_ms','whole_path_tflops','gpu_launches')}, d['roofline']['achieved'], d['e2e']['value'], d['effective']['value'], d['vae_decode']['ms_per_stage'], d['clocks'])
PY
