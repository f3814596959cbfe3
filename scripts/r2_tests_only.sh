mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 ${1:+-k "$1"} > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/r2_pytest.log | tail -30
