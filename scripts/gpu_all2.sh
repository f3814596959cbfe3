mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for c in cfg2 cfg3; do
timeout 600 python bench.py --steps 10 --warmup 3 --workload $c --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
python -c "
import json
d=json.load(open('gpurun_out/bench_$c.json'))
print('$c', round(d['ms_per_step'],3), 'ms/step', round(d['value']/1e6,2), 'M edge-layers/s; e2e', round(d['e2e']['value']/1e6,2), d['roofline']['kernel'], round(d['roofline']['us_per_launch'],1), 'us frac', round(d['roofline']['frac'],5), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})"
done
