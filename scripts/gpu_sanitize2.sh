mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/sanitize_memcheck_tests.log 2>&1; echo "memcheck rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" gpurun_out/sanitize_memcheck_tests.log | head -10
timeout 1500 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cfg5_knn30 or cfg4_nms20 or degenerate or switches or ragged" > gpurun_out/sanitize_racecheck_tests.log 2>&1; echo "racecheck rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_racecheck_tests.log | head -10
