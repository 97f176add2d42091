#!/bin/bash
# A/B of two builds of the product library on the SAME box (box-to-box variance is +-3%, more than most single changes):
#   mkdir pl-nerf_b200/ab; build the two variants into pl-nerf_b200/ab/{old,new}.so (git-ignored, they travel with gpurun);
#   gpurun -- 'bash tests/gpu_ab.sh old.so new.so'
# alternates the two files into place, three rounds, and prints the TrainStep time of each run
set -e
cd "$(dirname "$0")/.."
cp pl-nerf_b200/libplnerf_b200.so /tmp/keep.so
for i in 1 2 3; do
  for v in "$1" "$2"; do
    cp "pl-nerf_b200/ab/$v" pl-nerf_b200/libplnerf_b200.so
    echo -n "$v: "; PLNERF_ITERS=${PLNERF_ITERS:-600} python tests/gpu_train_step_target.py | python -c "import json,sys; print(json.loads(sys.stdin.read())['device_ms_per_iter'])"
  done
done
cp /tmp/keep.so pl-nerf_b200/libplnerf_b200.so
