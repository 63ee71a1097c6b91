#!/bin/bash
# Summarise the `ncu --set full` reports a tools/gpu_profile.sh run left in gpurun_out/ (run HERE, no GPU needed):
#   tools/ncu_full_summary.sh r02f > profiles/r02f_ncu_full_summary.txt
TAG=${1:-r02}
echo "ncu --set full --clock-control none --import-source on (one launch each; tools/gpu_profile.sh $TAG); raw-page metrics + source lines with most stall samples"
for rep in gpurun_out/${TAG}_prof_*.ncu-rep; do
  c=$(basename $rep .ncu-rep); c=${c#${TAG}_prof_}
  echo; echo "=== kernel_bench case $c"
  python tools/ncu_src.py $rep 8 | grep -v "^columns:"
done
