# Developer helper (GPU): short bench runs over the pipeline knobs; prints value / e2e per setting.
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-solver --no-other-configs"
for args in "" "--launch-fields 40" "--depth 6" "--depth 4" "--launch-fields 40 --depth 4" "--warps-per-sm 0 --depth 3"; do
  $B $args 2> gpurun_out/sweep.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('%-34s value %.4e  e2e %.4e  ms/step %.2f  frac %.4f' % (sys.argv[1] or '(default)', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']))" "$args"
done
