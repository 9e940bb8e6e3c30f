# ncu --set full of ONE launch of the frame-path LM kernel at its full residency (12 warps per SM: what the SMs see in the
# pipelined step), 160 frames = ~810 000 fits.  Usage: bash tools/gpu_prof_lm12.sh <tag>
TAG=${1:-x}
FSQ_WARPS=0 python tools/gpu_fit_prof.py fast 2 160 > gpurun_out/prof_lm12_$TAG.plain.log 2>&1 || exit 1
FSQ_WARPS=0 ncu --set full --clock-control none --import-source on -k regex:lmwarp -s 1 -c 1 -o gpurun_out/prof_lm12_$TAG -f python tools/gpu_fit_prof.py fast 2 160 > gpurun_out/prof_lm12_$TAG.ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/prof_lm12_$TAG.plain.log
