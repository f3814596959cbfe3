# full GPU suite + cfg2/cfg4 bench lines (no CPU leg)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/r2_pytest.log | tail -30
for w in ${WORKLOADS:-cfg2 cfg4}; do
timeout 300 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/r2_quick_$w.json 2> gpurun_out/r2_quick_$w.err; echo "bench $w rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_quick_$w.json"))
print("$w", "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, "us/launch", round(d["roofline"]["us_per_launch"],1))
PY
done
