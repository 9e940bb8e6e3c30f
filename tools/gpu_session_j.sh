#!/bin/bash
OUT=gpurun_out
python tools/gpu_cons_prof.py 2>&1 | tee $OUT/cons_prof_j.log
ncu --set full --clock-control none --import-source on -k regex:"cons_union|cons_process" -c 2 -o $OUT/prof_cons_j -f python tools/gpu_cons_prof.py > $OUT/ncu_cons_j.log 2>&1
echo "ncu rc=$?"
