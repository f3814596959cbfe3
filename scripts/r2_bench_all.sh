# bench lines of every single-GPU workload except the default one (no CPU leg)
mkdir -p gpurun_out
for w in ${WORKLOADS:-cfg1 cfg3 cfg4 cfg5 cfg5d}; do
timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo "bench $w rc=$?"; tail -2 gpurun_out/r2_bench_$w.err | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_$w.json"))
    print("$w", "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "M/s", round(d["value"]/1e6,2), {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, "us/launch", round(d["roofline"]["us_per_launch"],1))
except Exception as e: print("$w", "no line", e)
PY
done
