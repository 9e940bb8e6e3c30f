#!/bin/bash
# GPU-box session (round 1f): claim-ahead queue + finish kernel; occupancy variants of the FAST kernel under the bench.
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q -rA -x > $OUT/pytest_gpu_r01f.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_r01f.log
tail -4 $OUT/pytest_gpu_r01f.log
short() { python - "$1" <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        j=json.loads(ln); print("  value %.4g e2e %.4g ms/step %.3f fit_ms %s frac %.4f serial %s" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["ms_per_launch"], j["roofline"]["frac"], j.get("serial_ms_per_step")))
PY
}
for mb in 2 3 4; do
  bash tools/gpu_variant.sh "-DWMINB=$mb" f_mb$mb
  for cfg in "4 6" "4 8" "8 3" "0 3"; do
    set -- $cfg
    python bench.py --steps 200 --no-cpu-baseline --no-parity-solver --warps-per-sm $1 --depth $2 > $OUT/bench_f_mb${mb}_w$1_d$2.json 2> $OUT/bench_f_mb${mb}_w$1_d$2.err
    echo "mb=$mb wps=$1 depth=$2 rc=$?"; short $OUT/bench_f_mb${mb}_w$1_d$2.json
  done
done
