# stage timing: rebuild with the clock64 stamps compiled in (scratch copy on the GPU box), run the stamps script
mkdir -p gpurun_out
GCPNET_NVCC_FLAGS=-DGCP_STAMPS=1 python -m gcpnet_b200.build --force > gpurun_out/r2_stamps_build.log 2>&1 || tail -5 gpurun_out/r2_stamps_build.log
timeout 300 python scripts/tc_bwd_stamps.py > gpurun_out/r2_stamps_cfg2.log 2>&1; cat gpurun_out/r2_stamps_cfg2.log | tail -30
timeout 300 python scripts/tc_bwd_stamps.py 128 20 > gpurun_out/r2_stamps_cfg4.log 2>&1; cat gpurun_out/r2_stamps_cfg4.log | grep "bwd GCP3\|fwd GCP3"
