mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default rc=$? lines=$(wc -l < gpurun_out/bench_default.json)"
timeout 300 python bench.py --no-graph --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; echo "nograph rc=$? lines=$(wc -l < gpurun_out/bench_nograph.json)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu rc=$? lines=$(wc -l < gpurun_out/bench_2gpu.json)"
python - <<'PY'
import json
for f in ("default", "nograph", "2gpu"):
    d = json.load(open(f"gpurun_out/bench_{f}.json"))
    print(f, round(d["ms_per_step"], 4), "ms/step", round(d["value"] / 1e6, 2), "M/s; e2e", round(d["e2e"]["value"] / 1e6, 2), d["e2e"]["h2d_bytes_per_step"], d["n_gpus"])
PY
