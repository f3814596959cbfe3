mkdir -p gpurun_out
for rep in 1 2; do
for m in 0 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --option post_fused=$m > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json
d=json.load(open('gpurun_out/ab.json'))
print('post mode $m', round(d['ms_per_step'],4), 'ms/step; e2e', round(d['e2e']['ms_per_step'],4))"
done
done
