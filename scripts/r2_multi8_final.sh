# 8-GPU lines (peer-memory transport) + one-GPU lines of the same workloads on the same box
N=${1:-8}
mkdir -p gpurun_out
for w in cfg2 cfg4s; do
PORT=$((29511 + RANDOM % 200)); GCPNET_BENCH_TIMEOUT=100 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/r2_scale_${w}_n${N}_p2p.json 2> gpurun_out/r2_scale_${w}_n${N}_p2p.err; echo "bench $w N=$N rc=$?"
timeout 120 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/r2_scale_${w}_n1.json 2> gpurun_out/r2_scale_${w}_n1.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale_${w}_n${N}_p2p.json")); d1=json.load(open("gpurun_out/r2_scale_${w}_n1.json"))
    print("$w N=$N", "ms/step", round(d["ms_per_step"],4), "M/s", round(d["value"]/1e6,2), "| N=1 ms/step", round(d1["ms_per_step"],4), "M/s", round(d1["value"]/1e6,2), "| ratio", round(d["value"]/d1["value"],3), d["method"]["gradient_exchange"][:40])
except Exception as e: print("no line", e)
PY
done
