#!/bin/bash
# Developer sweep (GPU box): bench.py over (stacks per launch, depth, warps per SM, park_after) -> one line each.
#   bash tools/gpu_sweep.sh "G depth wps park" ...
OUT=gpurun_out
for cfg in "$@"; do
  set -- $cfg
  extra=""; [ "$4" != "" ] && [ "$4" != "0" ] && extra="--park $4"
  timeout 600 python bench.py --steps 320 --no-cpu-baseline --no-parity-solver --stacks-per-launch $1 --depth $2 --warps-per-sm $3 $extra > $OUT/bench_sw_g$1_d$2_w$3_p$4.json 2> $OUT/bench_sw_g$1_d$2_w$3_p$4.err
  python - $OUT/bench_sw_g$1_d$2_w$3_p$4.json "G=$1 depth=$2 wps=$3 park=$4" <<'PY'
import json,sys
ok=False
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        ok=True
        j=json.loads(ln); print("  %-30s value %.4g e2e %.4g ms/step %.3f fit_ms/launch %.3f pipe %.4f" % (sys.argv[2], j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["ms_per_launch"], j["roofline"]["frac_in_pipeline"]))
if not ok: print(sys.argv[2], "no JSON")
PY
done
