mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 600 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_b.log 2>&1
python - <<'PY'
import json
for c in ("cfg2","cfg4","cfg1"):
    d=json.load(open(f'gpurun_out/bench_{c}.json'))
    print(c, round(d['ms_per_step'],3), 'ms', round(d['value']/1e6,2), 'M/s; e2e', round(d['e2e']['value']/1e6,2), 'launches', d['gpu_launches'], d['roofline']['kernel'], round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],5), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})
print(open('gpurun_out/bench_ref.json').read()[:200])
PY
