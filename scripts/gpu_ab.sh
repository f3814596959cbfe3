mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
for opt in "post_fused=1" "post_fused=0"; do
  o=""; for kv in $opt; do o="$o --option $kv"; done
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $o > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json
d=json.load(open('gpurun_out/ab.json'))
print('$opt', round(d['ms_per_step'],4), 'ms/step; e2e', round(d['e2e']['ms_per_step'],4))"
done
done
