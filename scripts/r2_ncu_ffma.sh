mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_bwd_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_ffma_edge_bwd python bench.py --steps 1 --warmup 3 --workload cfg3 --no-cpu-baseline --no-graph > gpurun_out/r2_ncu_ffma.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
