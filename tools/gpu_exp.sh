set -u
OUT=gpurun_out/r2h; mkdir -p $OUT
run() { # name lib env
  env $3 LPHASH_B200_LIB="$2" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cfg5 > $OUT/bench_$1.json 2> $OUT/bench_$1.err
  python - "$1" "$OUT/bench_$1.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:22s} kernel_ms {d['roofline']['kernel_ms']:.4f}  ms/step {d['ms_per_step']:.4f} frac {d['roofline']['frac']:.3f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
D=$PWD/lphash_b200/liblphash_b200.so; N=$PWD/lphash_b200/liblphash_b200_nohint.so
run default $D "X=1"
run nowindow $D "LPHB_NO_L2_WINDOW=1"
run nohint $N "X=1"
run nohint_nowindow $N "LPHB_NO_L2_WINDOW=1"
run default2 $D "X=1"
