#!/bin/bash
# One gpurun call: launch lists (cold cache, with DRAM bytes) of the main and the grouped forward, ncu --set full of the
# kernel zoo, configs 3 / 4, SAM encoder and VAE timing.  Outputs under gpurun_out/.
mkdir -p gpurun_out
for w in main fused; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --nvtx --nvtx-include "unet_forward/" --csv --log-file gpurun_out/r02_launches_$w.csv python scripts/profile_fused.py $w > gpurun_out/prof_$w.log 2>&1
tail -1 gpurun_out/prof_$w.log
done
timeout 900 ncu --set full --import-source on --clock-control none --nvtx --nvtx-include "zoo/" -o gpurun_out/r02_zoo -f python scripts/kernel_zoo.py > gpurun_out/zoo.log 2>&1; tail -1 gpurun_out/zoo.log
python scripts/run_configs.py 3 4 > gpurun_out/r02_configs_3_4.jsonl 2> gpurun_out/r02_configs_3_4.err; cat gpurun_out/r02_configs_3_4.jsonl
python scripts/sam_bench.py > gpurun_out/r02_sam_bench.json 2> gpurun_out/r02_sam_bench.err; cat gpurun_out/r02_sam_bench.json
python scripts/vae_bench.py > gpurun_out/r02_vae_bench.json 2>&1; tail -1 gpurun_out/r02_vae_bench.json
