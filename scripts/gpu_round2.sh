#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -x -q --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 600 python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; grep -E "attn_self" gpurun_out/kernel_bench.log
timeout 1500 python bench.py --steps 2 --warmup 3 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','unet_step_ms','whole_path_tflops','gpu_launches','cpu_baseline')}, d['roofline']['achieved'], d['e2e']['value'])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json | cut -c1-400
