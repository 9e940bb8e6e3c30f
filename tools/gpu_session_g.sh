#!/bin/bash
# GPU-box session (round 1g): device consolidation + packed PSFs; bench through fetch=psfs; PCIe probe.
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q -rA > $OUT/pytest_gpu_r01g.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_r01g.log
grep -E "FAILED|ERROR|passed|failed" $OUT/pytest_gpu_r01g.log | tail -15
python tools/gpu_pcie.py 2>&1 | tee $OUT/pcie_r01g.log
short() { python - "$1" <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        j=json.loads(ln); print("  value %.4g e2e %.4g ms/step %.3f fit_ms %.3f frac %.4f serial %.3f d2h %d enqueue %.3f" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["ms_per_launch"], j["roofline"]["frac"], j["serial_ms_per_step"], j["e2e"]["d2h_bytes_per_step"], j["host_enqueue_ms_per_step"]))
PY
}
for cfg in "psfs 6" "candidates 6" "psfs 8" "psfs 4"; do
  set -- $cfg
  python bench.py --steps 200 --no-cpu-baseline --no-parity-solver --fetch $1 --depth $2 > $OUT/bench_g_$1_d$2.json 2> $OUT/bench_g_$1_d$2.err
  echo "fetch=$1 depth=$2 rc=$?"; short $OUT/bench_g_$1_d$2.json; tail -3 $OUT/bench_g_$1_d$2.err
done
python tools/gpu_tail_check.py 2>&1 | grep -E "maxiter 200 park  0|maxiter  15 park  0"
