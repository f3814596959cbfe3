mkdir -p gpurun_out
(timeout 300 python scripts/tc_bwd_check.py mid; timeout 300 python scripts/tc_bwd_check.py rand; timeout 300 python scripts/tc_bwd_stamps.py) > gpurun_out/tc_bwd.log 2>&1; echo "rc=$?" >> gpurun_out/tc_bwd.log
grep -E "tc=|over|bwd GCP[713]|rc=" gpurun_out/tc_bwd.log
