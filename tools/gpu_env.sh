#!/usr/bin/env bash
# kernel time under runtime switches / library variants, repeated.  Usage: tools/gpu_env.sh tag "ENV=1 ..|-" variant [repeat]
set -u
TAG="$1"; ENVS="$2"; V="$3"; N="${4:-2}"
OUT="gpurun_out/$TAG"; mkdir -p "$OUT"
lib="$PWD/lphash_b200/liblphash_b200_$V.so"; [ "$V" = default ] && lib="$PWD/lphash_b200/liblphash_b200.so"
[ "$ENVS" = "-" ] && ENVS=""
for i in $(seq 1 $N); do
  env $ENVS LPHASH_B200_LIB="$lib" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$OUT/b.json" 2> "$OUT/b.err"
  python -c "
import json;d=json.load(open('$OUT/b.json'));print('$V [$ENVS] kernel_ms %.4f'%d['roofline']['kernel_ms'])"
done
