#!/bin/bash
# GPU-box session (round 1e): full parity suite on the current build, then A/B variants of the FAST kernel.
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q -rA -x > $OUT/pytest_gpu_r01e.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_r01e.log
tail -15 $OUT/pytest_gpu_r01e.log
bash tools/gpu_variant.sh "-DWRECUR=0" e_norecur
bash tools/gpu_variant.sh "-DWRECUR=1" e_recur
bash tools/gpu_variant.sh "-DWRECUR=1 -DWLMPAR_MAX=2" e_recur_m2
bash tools/gpu_variant.sh "-DWRECUR=1 -DWLMPAR_MAX=3" e_recur_m3
