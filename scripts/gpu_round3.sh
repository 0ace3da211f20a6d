#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -25
timeout 600 python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; grep -E "pair|out1280_bn0|out640_bn0|ff2_1280_bn0|qkv1280_bn0|res64_bn0|res32_bn0|res128_bn0" gpurun_out/kernel_bench.log
for f in 1 0; do
  OMG_LN_FOLD=$f timeout 1500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_fold$f.err > gpurun_out/bench_fold$f.json; tail -1 gpurun_out/bench_fold$f.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_fold$f.json'))
print("OMG_LN_FOLD=$f", {k:d[k] for k in ('value','ms_per_step','unet_step_ms','whole_path_tflops','gpu_launches')}, d['roofline']['achieved'], d['e2e']['value'])
PY
done
