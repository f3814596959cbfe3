mkdir -p gpurun_out
for rep in 1 2; do
for pr in 0 -1; do
  GCPNET_MAIN_PRIORITY=$pr timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json
d=json.load(open('gpurun_out/ab.json'))
print('main priority $pr', round(d['ms_per_step'],4), 'ms/step; e2e', round(d['e2e']['ms_per_step'],4))"
done
done
