# A/B of the FAST LM kernel: baseline build (shipped) vs FSQ_WDEFS variants; each: gpu_lm_ab.py timing + a results checksum
set -e
python tools/gpu_lm_ab.py base > gpurun_out/lmab_base.log 2>&1 || true
for v in "$@"; do
  FSQ_WDEFS="$v" python -m fluorosequencingimageanalysis_b200.build --force > /dev/null 2>&1
  tag=$(echo "$v" | tr -c 'A-Za-z0-9\n' '_')
  FSQ_WDEFS="$v" python tools/gpu_lm_ab.py "$tag" > gpurun_out/lmab_$tag.log 2>&1 || true
done
tail -n 5 gpurun_out/lmab_*.log
