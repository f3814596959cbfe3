mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cfg2.err
cat gpurun_out/bench_cfg2.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-graph --no-cpu-baseline > gpurun_out/bench_cfg2_eager.json 2> gpurun_out/bench_cfg2_eager.err
cat gpurun_out/bench_cfg2_eager.json
