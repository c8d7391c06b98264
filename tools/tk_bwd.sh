#!/bin/bash
for wl in A1; do
  for v in "SPK_PULL_T=1" "SPK_PULL_T=1 SPK_PULL_THREADS=64" "SPK_PULL_T=1 SPK_PULL_THREADS=256" "SPK_PULL_T=2 SPK_PULL_THREADS=256" "SPK_PULL_T=4 SPK_PULL_THREADS=256 SPK_PULL_SMEM_KB=76"; do
    echo -n "$wl [$v] "; env $v python tools/time_kernels.py $wl 2>&1 | grep sp_gather_bwd
  done
done
