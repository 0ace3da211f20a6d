#!/bin/bash
mkdir -p gpurun_out
for lib in "" "$PWD/build/variants/libomg_nopf.so"; do
  echo "== lib [$lib]"
  OMG_B200_LIB=$lib timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_processors_gpu.py -m gpu -x -q --no-header -p no:cacheprovider -k "attention or processor or ip" 2>&1 | tail -3
  OMG_B200_LIB=$lib timeout 300 python scripts/attn_bench.py 2>&1 | grep -E "attn"
done
