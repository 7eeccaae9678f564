#!/usr/bin/env bash
# One gpurun call: GPU parity tests, smoke, bench line, ncu launch list, ncu full capture of the
# dominant kernel.  Everything lands under gpurun_out/<tag>/.   Usage: tools/gpu_round.sh <tag> [what...]
# what: tests smoke bench launches full (default: all)
set -u
TAG="${1:-r01}"; shift || true
WHAT="${*:-tests smoke bench launches full}"
OUT="gpurun_out/$TAG"
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,memory.total --format=csv > "$OUT/gpu.csv" 2>&1
has() { [[ " $WHAT " == *" $1 "* ]]; }
if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -5 "$OUT/pytest_gpu.log"
fi
if has smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
  echo "smoke exit $?" >> "$OUT/smoke.log"; tail -4 "$OUT/smoke.log"
fi
if has bench; then
  timeout 900 python bench.py --steps 20 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"
  echo "bench exit $?"; cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
fi
if has refbench; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
  echo "refbench exit $?"; cat "$OUT/bench_ref.json"
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file "$OUT/launches.csv" python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$OUT/launches_bench.log" 2>&1
  echo "launches exit $?"
fi
if has full; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_query_ -s 3 -c 1 \
    -f -o "$OUT/prof_query_tiled" python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$OUT/full_bench.log" 2>&1
  echo "full exit $?"; ls -la "$OUT"
fi
