#!/bin/bash
for wl in A N8192; do
  for v in "SPK_TC_SPLIT=1" "SPK_TC_SPLIT=2" "SPK_X=0" "SPK_TC_SPLIT=4"; do
    echo -n "$wl [$v] "; env $v python tools/time_kernels.py $wl > gpurun_out/tmp.txt 2>&1; grep chamfer_fwd gpurun_out/tmp.txt
  done
done
