#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python scripts/run_configs.py 3 4 2>&1 | grep -E "^\{" > gpurun_out/configs_3_4.jsonl; cat gpurun_out/configs_3_4.jsonl | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
