# ncu --set full of the FFMA edge backward and of the off-tile weight-gradient product at cfg3 ((100,16) dims)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^edge_bwd_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_ffma_edge_bwd python bench.py --steps 1 --warmup 3 --workload cfg3 --no-cpu-baseline --no-graph > gpurun_out/r2_ncu_ffma.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:node_wgrad_kernel -s 7 -c 1 -f -o gpurun_out/r2_prof_edge_wgrad python bench.py --steps 1 --warmup 3 --workload cfg3 --no-cpu-baseline --no-graph > gpurun_out/r2_ncu_wgrad.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
