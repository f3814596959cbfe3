# compute-sanitizer memcheck over the tests that exercise the round-2 kernels added last (in-tile segment sums with hub
# destinations, GCP-Baseline switches, message attention, off-tile weight gradients with row chunks, the CPD model)
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q --timeout=400 -x \
  -k "segment_sums or interactions2 or baseline_variants or cfg3_six or message_passing_alone or cpd_model or tensor_core_and_ffma" \
  > gpurun_out/r2_sanitize.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/r2_sanitize.log
grep -E "ERROR SUMMARY|passed|failed|rc=|Invalid|out of bounds" gpurun_out/r2_sanitize.log | tail -12
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q --timeout=400 -x \
  -k "segment_sums or eq_layer2 or tiny_no_vector_gate" > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_racecheck.log
grep -E "RACECHECK SUMMARY|passed|failed|rc=|hazard" gpurun_out/r2_racecheck.log | tail -8
