mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool initcheck --print-limit 100000 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/initcheck.log 2>&1
grep -E "Uninitialized|at .*\+0x|by thread" gpurun_out/initcheck.log | grep -E " at " | sed -E 's/\+0x[0-9a-f]+.*//' | sort | uniq -c | sort -rn | head -12
grep -c "Uninitialized" gpurun_out/initcheck.log
