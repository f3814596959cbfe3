mkdir -p gpurun_out
export GCPNET_MAIN_PRIORITY=0
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -6 gpurun_out/sanitize_racecheck.log
