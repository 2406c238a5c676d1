#!/bin/bash
# usage: tools/gpu_r02r.sh : index-builder parity on the GPU + builder timing (device pipeline and host walk)
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_index_build.py -m gpu -q ) 2>&1 | tail -25
timeout 600 python tools/build_probe.py 16 2>&1 | tail -5 | tee gpurun_out/build_probe_r02r.txt
PM_BUILD_HOST_WALK=1 timeout 600 python tools/build_probe.py 0 2>&1 | tail -5 | tee gpurun_out/build_probe_r02r_hostwalk.txt
