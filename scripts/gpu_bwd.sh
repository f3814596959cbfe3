mkdir -p gpurun_out
timeout 300 python scripts/tc_bwd_check.py test51 > gpurun_out/tc_bwd.log 2>&1; echo "rc=$?" >> gpurun_out/tc_bwd.log
cat gpurun_out/tc_bwd.log
