mkdir -p gpurun_out
timeout 300 python scripts/tc_check.py all > gpurun_out/tc_check.log 2>&1; echo "rc=$?" >> gpurun_out/tc_check.log
tail -30 gpurun_out/tc_check.log
