# N-GPU job: NCCL gradient-averaging test + weak (cfg2) and strong (cfg4s) scaling lines
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ddp.py -m gpu -q --timeout=120 2>&1 | tail -5
for w in cfg2 cfg4s; do
PORT=$((29511 + RANDOM % 200)); GCPNET_BENCH_TIMEOUT=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 5 --workload $w > gpurun_out/r2_scale_${w}_n$N.json 2> gpurun_out/r2_scale_${w}_n$N.err; echo "bench $w N=$N rc=$?"
grep -i "captured all-reduce\|error\|Traceback" gpurun_out/r2_scale_${w}_n$N.err | head -5
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale_${w}_n$N.json"))
    print("$w N=$N", "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "M/s", round(d["value"]/1e6,2), d["method"]["gradient_exchange"])
except Exception as e: print("no line", e)
PY
done
# one-GPU lines of the same workloads on the same box for the ratio
for w in cfg2 cfg4s; do
timeout 200 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/r2_scale_${w}_n1.json 2> gpurun_out/r2_scale_${w}_n1.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale_${w}_n1.json"))
    print("$w N=1", "ms/step", round(d["ms_per_step"],4), "M/s", round(d["value"]/1e6,2))
except Exception as e: print("no line", e)
PY
done
