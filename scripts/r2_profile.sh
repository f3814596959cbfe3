# round 2 profiles: launch list of cfg2 steps (eager launches under ncu), full captures of the four critical-path kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_ncu_list.log 2>&1
for k in tc_edge_bwd_kernel tc_edge_fwd_kernel node_bwd_kernel node_fwd_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/r2_prof_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_ncu_$k.log 2>&1
done
ls -la gpurun_out/r2_prof_*.ncu-rep gpurun_out/r2_launches_cfg2.csv
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"
cut -c1-400 gpurun_out/r2_bench_default.json; echo; cut -c1-300 gpurun_out/r2_bench_reference.json
