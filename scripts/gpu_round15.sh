#!/bin/bash
echo "== default"; timeout 300 python scripts/attn_bench.py 2>&1 | grep -E "attn"
echo "== TILES=1 (3 stages)"; OMG_ATTN_TILES=1 timeout 300 python scripts/attn_bench.py 2>&1 | grep -E "self"
echo "== TILES=1 (5 stages)"; OMG_ATTN_TILES=1 OMG_B200_LIB=$PWD/build/variants/libomg_s5.so timeout 300 python scripts/attn_bench.py 2>&1 | grep -E "attn"
echo "== G=2 8 stages"; OMG_B200_LIB=$PWD/build/variants/libomg_s8.so timeout 300 python scripts/attn_bench.py 2>&1 | grep -E "self"
