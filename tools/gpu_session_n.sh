#!/bin/bash
OUT=gpurun_out
python -m pytest tests/test_gpu_fast.py -m gpu -q -x 2>&1 | tail -3
for v in 0 1; do
  bash tools/gpu_variant.sh "-DWPOOL=$v" n_pool$v
  for f in psfs candidates; do
    python bench.py --steps 300 --no-cpu-baseline --no-parity-solver --fetch $f > $OUT/bench_n_pool${v}_$f.json 2> $OUT/bench_n_pool${v}_$f.err
    python - $OUT/bench_n_pool${v}_$f.json "pool=$v $f" <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        j=json.loads(ln); print("  %-20s value %.4g e2e %.4g ms/step %.3f fit_ms %.3f serial %.3f" % (sys.argv[2], j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["ms_per_launch"], j["serial_ms_per_step"]))
PY
  done
done
