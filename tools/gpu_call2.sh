#!/usr/bin/env bash
set -u
OUT=gpurun_out/r01d; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "chunk_pipeline or query_batch" 2>&1 | tail -3
bash tools/gpu_variants.sh r01d default p1 p1b7 p1b8 p2b7 p2nh p2b5
LPHB_NO_L2_WINDOW=1 bash tools/gpu_variants.sh r01d p2nh
grep -o '"e2e": {[^}]*}' $OUT/bench_default.json
