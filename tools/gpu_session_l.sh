#!/bin/bash
OUT=gpurun_out
for cfg in "4 6" "2 8" "2 10" "1 14" "1 18"; do
  set -- $cfg
  for f in psfs candidates; do
    python bench.py --steps 300 --no-cpu-baseline --no-parity-solver --fetch $f --warps-per-sm $1 --depth $2 > $OUT/bench_l_$f_w$1_d$2.json 2> $OUT/bench_l_$f_w$1_d$2.err
    python - $OUT/bench_l_$f_w$1_d$2.json "$f wps=$1 depth=$2" <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        j=json.loads(ln); print("  %-28s value %.4g e2e %.4g ms/step %.3f enqueue %.3f" % (sys.argv[2], j["value"], j["e2e"]["value"], j["ms_per_step"], j["host_enqueue_ms_per_step"]))
PY
  done
done
