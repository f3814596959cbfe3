mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:node_bwd_kernel -s 6 -c 1 -o gpurun_out/prof_node_bwd2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_node.log 2>&1
ls -la gpurun_out/*.ncu-rep
