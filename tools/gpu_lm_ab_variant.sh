# A/B of one FSQ_WDEFS variant of the FAST LM kernel against the shipped build: timing + hash, then the parity tests on
# the variant build.  Usage: bash tools/gpu_lm_ab_variant.sh "<nvcc defines>"
V="$1"
python tools/gpu_lm_ab.py base 3 | cut -c1-220
export FSQ_WDEFS="$V"
python -m fluorosequencingimageanalysis_b200.build --force > /dev/null 2>&1
python tools/gpu_lm_ab.py variant 3 | cut -c1-220
python -m pytest tests/test_gpu_fast.py tests/test_gpu_fit.py tests/test_gpu_parity_table.py -x -q -s 2>&1 | grep "parity\[\|chi2\[\|passed\|failed\|Error" | cut -c1-200
python tools/gpu_fit11_variants.py 200000 0 2>&1 | grep "float64" | grep "thread\|hybrid8" | cut -c1-100
