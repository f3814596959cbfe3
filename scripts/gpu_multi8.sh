mkdir -p gpurun_out
N=${NG:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
wc -l gpurun_out/bench_${N}gpu.json
python -c "
import json
d=json.load(open('gpurun_out/bench_${N}gpu.json'))
print('${N}gpu', round(d['ms_per_step'],4), 'ms/step', round(d['value']/1e6,2), 'M/s; e2e', round(d['e2e']['ms_per_step'],4), d['clocks'])"
grep -v -i "warn" gpurun_out/bench_${N}gpu.err | tail -3 | cut -c1-200
