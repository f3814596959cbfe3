# final build: full GPU suite, smoke(), bench lines of every workload (default line with the CPU leg), reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/r2_pytest.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
WORKLOADS="cfg1 cfg3 cfg4 cfg5 cfg5d" bash scripts/r2_bench_all.sh
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench default rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_default.json")); r=json.load(open("gpurun_out/r2_bench_reference.json"))
print("default", round(d["ms_per_step"],4), "ms", round(d["value"]/1e6,2), "M/s e2e", round(d["e2e"]["value"]/1e6,2), "cpu", round(d["cpu_baseline"]["value"]/1e6,3), "ref arm", round(r["value"]/1e6,3), "launches", d["gpu_launches"], d["clocks"])
PY
