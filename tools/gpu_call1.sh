#!/usr/bin/env bash
# session call 1: validate state + tuning variants
set -u
OUT=gpurun_out/r01c; mkdir -p $OUT
bash tools/gpu_round.sh r01c tests smoke bench
bash tools/gpu_variants.sh r01c default nohint p6 p2 b8 b5 b4p6
export LPHB_BENCH_NOCHECK=1
bash tools/gpu_variants.sh r01c nogather
unset LPHB_BENCH_NOCHECK
LPHB_NO_L2_WINDOW=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_nowindow.json 2> $OUT/bench_nowindow.err
python -c "
import json;d=json.load(open('$OUT/bench_nowindow.json'));print('nowindow kernel_ms',d['roofline']['kernel_ms'])"
LPHB_NO_L2_WINDOW=1 LPHASH_B200_LIB=$PWD/lphash_b200/liblphash_b200_nohint.so timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_nowindow_nohint.json 2> $OUT/bench_nowindow_nohint.err
python -c "
import json;d=json.load(open('$OUT/bench_nowindow_nohint.json'));print('nowindow+nohint kernel_ms',d['roofline']['kernel_ms'])"
