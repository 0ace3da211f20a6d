#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_pipeline_gpu.py -x -q --no-header -p no:cacheprovider 2>&1 | tail -4
for mode in 0 1; do
  OMG_NO_PDL=$mode timeout 1500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_pdl$mode.err > gpurun_out/bench_pdl$mode.json; tail -1 gpurun_out/bench_pdl$mode.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_pdl$mode.json'))
print("OMG_NO_PDL=$mode", {k:d[k] for k in ('value','ms_per_step','unet_step_ms','whole_path_tflops')}, d['roofline']['achieved'], d['e2e']['value'])
PY
done
