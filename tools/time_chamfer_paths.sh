#!/bin/bash
# dense vs sorted Chamfer forward over a grid of shapes (device time, CUDA-graph replay): bash tools/time_chamfer_paths.sh
for s in "32 1024 1024" "32 2048 2048" "32 4096 4096" "32 8192 8192" "8 16384 16384" "1 2048 16384" "1 4096 16384" "4 4096 2048" "8 2048 2048"; do
  for p in dense sorted; do
    echo -n "$p: "; SPK_CHAMFER_PATH=$p timeout 120 python tools/time_chamfer.py $s
  done
done
