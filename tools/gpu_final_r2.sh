#!/usr/bin/env bash
# Round-2 evidence run: launch list of the default bench command, one ncu --set full capture of the query kernel
# (steady state), the bench lines themselves (ours + reference arm), rows.  Usage: tools/gpu_final_r2.sh <tag>
set -u
TAG="$1"; OUT="gpurun_out/$TAG"; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > "$OUT/gpu.csv" 2>&1
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "reference rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cfg5 > "$OUT/launches.log" 2>&1; echo "launch list rc $?"
timeout 1200 ncu --set full --clock-control none --cache-control none --import-source on \
  -k regex:k_query_tiled -s 8 -c 1 -f -o "$OUT/prof_query" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-cfg5 > "$OUT/prof_query.log" 2>&1; echo "prof rc $?"
timeout 900 python tools/bench_rows.py scan part3 > "$OUT/rows.jsonl" 2> "$OUT/rows.err"; echo "rows rc $?"
ls -la "$OUT"
