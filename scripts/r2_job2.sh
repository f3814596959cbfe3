# round 2, job 2: full GPU suite + bench lines of every workload (no CPU leg except cfg2)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/r2_pytest.log | tail -12
for w in cfg1 cfg2 cfg3 cfg4 cfg5 cfg5d; do
timeout 900 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo "bench $w rc=$?"; tail -2 gpurun_out/r2_bench_$w.err | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_$w.json"))
    print("$w", "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "M/s", round(d["value"]/1e6,2), {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, "us/launch", round(d["roofline"]["us_per_launch"],1))
except Exception as e: print("$w", "no line", e)
PY
done
