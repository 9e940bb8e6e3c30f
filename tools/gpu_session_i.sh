#!/bin/bash
# GPU-box session (round 1i): batched consolidation scans -- tests, launch list, bench both fetch modes, lmwarp ncu capture.
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q -rA > $OUT/pytest_gpu_r01i.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_r01i.log
grep -E "FAILED|passed|failed" $OUT/pytest_gpu_r01i.log | tail -8
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-solver"
$BCMD > $OUT/plain_r01i.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_r01i.csv $BCMD > $OUT/ncu_launch_r01i.log 2>&1
echo "ncu launches rc=$?"
for f in psfs candidates; do
  python bench.py --steps 300 --no-cpu-baseline --no-parity-solver --fetch $f > $OUT/bench_i_$f.json 2> $OUT/bench_i_$f.err
  python - $OUT/bench_i_$f.json <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        j=json.loads(ln); print("  %s value %.4g e2e %.4g ms/step %.3f fit_ms %.3f frac %.4f serial %.3f d2h %d" % (sys.argv[1], j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["ms_per_launch"], j["roofline"]["frac"], j["serial_ms_per_step"], j["e2e"]["d2h_bytes_per_step"]))
PY
done
$BCMD > $OUT/plain2_r01i.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lmwarp -s 3 -c 1 -o $OUT/prof_lmwarp_i -f $BCMD > $OUT/ncu_full_r01i.log 2>&1
echo "ncu full lmwarp rc=$?"
