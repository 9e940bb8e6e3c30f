#!/bin/bash
# One GPU-box session: parity tests, bench (both arms), ncu launch list + full captures of the top kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.log 2>&1
python -m pytest tests -m gpu -q -rA > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python bench.py --fetch candidates --no-cpu-baseline --no-parity-solver > $OUT/bench_cand_$TAG.json 2> $OUT/bench_cand_$TAG.err; echo "bench candidates rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?"
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-solver"
$BCMD > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $BCMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
$BCMD > $OUT/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lmwarp -s 3 -c 1 -o $OUT/prof_lmwarp_$TAG -f $BCMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full lmwarp rc=$?"
$BCMD > $OUT/plain3_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:detect_cm -s 3 -c 1 -o $OUT/prof_detect_$TAG -f $BCMD > $OUT/ncu_full_det_$TAG.log 2>&1
echo "ncu full detect rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"cons_adj|cons_component" -s 4 -c 2 -o $OUT/prof_cons_$TAG -f $BCMD > $OUT/ncu_full_cons_$TAG.log 2>&1
echo "ncu full cons rc=$?"
tail -3 $OUT/pytest_gpu_$TAG.log
cat $OUT/bench_$TAG.json
