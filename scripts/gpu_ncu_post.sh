mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_post_fused_kernel -s 6 -c 1 -f -o gpurun_out/prof_post python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_post.log 2>&1
ls -la gpurun_out/prof_post.ncu-rep
