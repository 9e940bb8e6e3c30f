#!/bin/bash
# GPU-box session (round 1h): launch list of the new pipeline (consolidation cost), damped-first lmpar variant.
OUT=gpurun_out
mkdir -p $OUT
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-solver"
$BCMD > $OUT/plain_r01h.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_r01h.csv $BCMD > $OUT/ncu_launch_r01h.log 2>&1
echo "ncu launches rc=$?"
bash tools/gpu_variant.sh "-DWPARFIRST=0" h_base
bash tools/gpu_variant.sh "-DWPARFIRST=1" h_parfirst
python bench.py --steps 200 --no-cpu-baseline --no-parity-solver --fetch candidates > $OUT/bench_h_parfirst.json 2> $OUT/bench_h_parfirst.err
python - <<'PY'
import json
for ln in open("gpurun_out/bench_h_parfirst.json"):
    if ln.startswith("{"):
        j=json.loads(ln); print("parfirst bench (candidates): value %.4g e2e %.4g ms/step %.3f fit_ms %.3f" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["ms_per_launch"]))
PY
