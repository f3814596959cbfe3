mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_edge_fwd_kernel -s 6 -c 1 -o gpurun_out/prof_tc_edge_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_tcfwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_edge_bwd_kernel -s 6 -c 1 -o gpurun_out/prof_tc_edge_bwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_tcbwd.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_b.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 600 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ls -la gpurun_out/*.ncu-rep gpurun_out/*.json
