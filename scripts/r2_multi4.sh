# 4-GPU line of the default workload (the driver's scaling run uses N = 1, 2, 4, 8)
mkdir -p gpurun_out
PORT=$((29511 + RANDOM % 200)); GCPNET_BENCH_TIMEOUT=90 timeout 130 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_scale_cfg2_n4_p2p.json 2> gpurun_out/r2_scale_cfg2_n4_p2p.err; echo "bench N=4 rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale_cfg2_n4_p2p.json"))
    print("cfg2 N=4", "ms/step", round(d["ms_per_step"],4), "M/s", round(d["value"]/1e6,2), d["method"]["gradient_exchange"][:50], "cpu" in d and d.get("cpu_baseline"))
except Exception as e: print("no line", e)
PY
tail -3 gpurun_out/r2_scale_cfg2_n4_p2p.err | cut -c1-200
