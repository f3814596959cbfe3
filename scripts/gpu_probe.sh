set -x
mkdir -p gpurun_out
cd tests/umma
for c in ts tsfast timing; do
  PROBE_DUMP=1 timeout 60 ./umma_probe $c >> ../../gpurun_out/probe.log 2>&1; echo "case $c rc=$?" >> ../../gpurun_out/probe.log
done
cat ../../gpurun_out/probe.log
