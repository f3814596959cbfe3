mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
bash scripts/gpu_ncu_list.sh
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cfg2.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/bench_cfg2.json'))
print('cfg2', d['ms_per_step'], d['value'], d['e2e'], d['roofline']['kernel_ms_per_step'], d['cpu_baseline'])"
