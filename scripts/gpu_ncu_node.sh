mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"node_(bwd|fwd)_kernel" -s 12 -c 2 -f -o gpurun_out/prof_node3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_node.log 2>&1
ls -la gpurun_out/*.ncu-rep
