#!/bin/bash
for wl in A A1 N8192; do
  for v in "SPK_X=0" "SPK_FWD_THREADS=128" "SPK_FWD_THREADS=128 SPK_FWD_TILE_KB=32" "SPK_FWD_THREADS=256 SPK_FWD_TILE_KB=16" "SPK_FWD_THREADS=128 SPK_FWD_TILE_KB=8"; do
    echo -n "$wl [$v] "; env $v python tools/time_kernels.py $wl > gpurun_out/tmp.txt 2>&1; grep sp_gather_fwd gpurun_out/tmp.txt
  done
done
