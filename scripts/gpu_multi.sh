mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"
wc -l gpurun_out/bench_2gpu.json
python -c "
import json
d=json.load(open('gpurun_out/bench_2gpu.json'))
print('2gpu', round(d['ms_per_step'],4), 'ms/step', round(d['value']/1e6,2), 'M/s; e2e', round(d['e2e']['ms_per_step'],4))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; echo "rc=$?"
wc -l gpurun_out/bench_2gpu_ref.json; head -c 200 gpurun_out/bench_2gpu_ref.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_1gpu.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/bench_1gpu.json'))
print('1gpu same box', round(d['ms_per_step'],4), 'ms/step')"
