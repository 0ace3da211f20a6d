#!/bin/bash
# Run every GPU kernel test function in its own process (a trapped kernel poisons the CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/unit_gpu.txt 2>&1
python - <<'PY' > gpurun_out/unit_tests.txt
import subprocess, re
out = subprocess.run(["python", "-m", "pytest", "tests/test_kernels_gpu.py", "--collect-only", "-q"], capture_output=True, text=True).stdout
funcs = []
for line in out.splitlines():
    m = re.match(r"(tests/test_kernels_gpu.py::\w+)", line)
    if m and m.group(1) not in funcs:
        funcs.append(m.group(1))
print("\n".join(funcs))
PY
: > gpurun_out/unit.log
for t in $(cat gpurun_out/unit_tests.txt); do
  echo "=== $t" >> gpurun_out/unit.log
  timeout 300 python -m pytest "$t" -q -x --no-header -p no:cacheprovider 2>&1 | tail -25 >> gpurun_out/unit.log
  echo "exit=$?" >> gpurun_out/unit.log
done
grep -E "^===|passed|failed|error|Error|timeout|exit=" gpurun_out/unit.log | tail -80
