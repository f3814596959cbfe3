mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"
grep -v -i "warn" gpurun_out/bench_2gpu.err | tail -25 | cut -c1-400
cat gpurun_out/bench_2gpu.json | cut -c1-600
