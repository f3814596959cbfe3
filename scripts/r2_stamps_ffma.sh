mkdir -p gpurun_out
GCPNET_NVCC_FLAGS=-DGCP_STAMPS=1 python -m gcpnet_b200.build --force > gpurun_out/r2_stamps_build.log 2>&1 || tail -5 gpurun_out/r2_stamps_build.log
timeout 300 python scripts/ffma_bwd_stamps.py > gpurun_out/r2_stamps_ffma.log 2>&1; tail -20 gpurun_out/r2_stamps_ffma.log
