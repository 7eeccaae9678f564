#!/usr/bin/env bash
# Round-end measurement, one gpurun call: GPU parity tests, smoke, bench line (ours + reference arm),
# the other SURVEY 8 rows, ncu launch list and one full capture of the dominant kernel.
# Everything lands under gpurun_out/<tag>/.   Usage: tools/gpu_final.sh <tag>
set -u
TAG="${1:-final}"
OUT="gpurun_out/$TAG"; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,memory.total --format=csv > "$OUT/gpu.csv" 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -3 "$OUT/pytest_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke exit $?" >> "$OUT/smoke.log"; tail -3 "$OUT/smoke.log"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "refbench exit $?"
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?"; cat "$OUT/bench.json"
timeout 1200 python tools/bench_rows.py scan k63 reads > "$OUT/rows.jsonl" 2> "$OUT/rows.err"; echo "rows exit $?"; cut -c1-260 "$OUT/rows.jsonl"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$OUT/launches_bench.log" 2>&1; echo "launches exit $?"
timeout 1200 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_query_tiled -s 8 -c 1 \
  -f -o "$OUT/prof_query_tiled" python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/full_bench.log" 2>&1; echo "full exit $?"
ls -la "$OUT"
