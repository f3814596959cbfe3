#!/bin/bash
# build, then run a command on the GPU box; refuses to go if the build failed
cd /root/repo
python -m gcpnet_b200.build > /tmp/build.log 2>&1 || { grep -E "error" /tmp/build.log | head; echo "BUILD FAILED"; exit 1; }
/usr/local/graft/bin/gpurun --timeout ${TMO:-900} -- "$@"
