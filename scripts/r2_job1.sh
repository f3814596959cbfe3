# round 2, job 1: full GPU suite + cfg2 bench (default invocation) + launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2j1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j1_pytest.log
tail -25 gpurun_out/r2j1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j1_bench_cfg2.json 2> gpurun_out/r2j1_bench_cfg2.err; echo "bench rc=$?"
cat gpurun_out/r2j1_bench_cfg2.json
tail -5 gpurun_out/r2j1_bench_cfg2.err
