mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_edge_fwd_kernel -s 6 -c 1 -f -o gpurun_out/prof_tc_edge_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_tcfwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_edge_bwd_kernel -s 6 -c 1 -f -o gpurun_out/prof_tc_edge_bwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_tcbwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:node_fwd_kernel -s 6 -c 1 -f -o gpurun_out/prof_node_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_nodefwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:node_bwd_kernel -s 6 -c 1 -f -o gpurun_out/prof_node_bwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_nodebwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
bash scripts/gpu_final.sh
