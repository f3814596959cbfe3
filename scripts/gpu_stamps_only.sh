mkdir -p gpurun_out
timeout 300 python scripts/tc_bwd_stamps.py > gpurun_out/stamps.log 2>&1; tail -3 gpurun_out/stamps.log
