N=${1:-8}
mkdir -p gpurun_out
for t in p2p nccl; do
for w in cfg2 cfg4s; do
PORT=$((29511 + RANDOM % 200)); GCPNET_DDP_TRANSPORT=$t GCPNET_BENCH_TIMEOUT=120 timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 5 --workload $w > gpurun_out/r2_scale_${w}_n${N}_$t.json 2> gpurun_out/r2_scale_${w}_n${N}_$t.err; echo "bench $w N=$N $t rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale_${w}_n${N}_$t.json"))
    print("$w N=$N $t", "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "M/s", round(d["value"]/1e6,2))
except Exception as e: print("no line", e)
PY
done
done
