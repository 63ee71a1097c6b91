#!/bin/bash
# round 2, call 40: profile pass on the current code (kernel bench in graph, launch list, ncu --set full captures)
set -u
mkdir -p gpurun_out
bash tools/gpu_profile.sh r02al > gpurun_out/r02al_profile_script.log 2>&1
tail -25 gpurun_out/r02al_profile_script.log
