#!/bin/bash
# Developer A/B helper (GPU box): rebuild libfsq.so with extra -D switches for the FAST kernel and run the checks.
#   tools/gpu_variant.sh "<defs>" <tag>
set -e
FSQ_WDEFS="$1" python -m fluorosequencingimageanalysis_b200.build > /dev/null
python tools/gpu_fast_check.py 2>&1 | grep -E "^fast  |fast  +faithful|^fast   " > gpurun_out/variant_$2.log
python tools/gpu_tail_check.py 2>&1 | grep -E "maxiter 200 park  0|maxiter  15 park  0" >> gpurun_out/variant_$2.log
echo "== $2 ($1)"; cat gpurun_out/variant_$2.log
