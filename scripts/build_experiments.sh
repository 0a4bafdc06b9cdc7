#!/bin/bash
# Build the stand-alone tuner once per experimental kernel variant into bin/ (git-ignored, travels
# to the GPU box with the repo snapshot):
#   bin/tune_v0  library code path        bin/tune_v1  stage first (residual/velocity)
#   bin/tune_v2  + column clusters (CLn)  bin/tune_v4  register prefetch two planes deep
# then, on a GPU box, e.g.
#   for v in 0 1 2 4; do bin/tune_v$v 2d 1024 2d 4096 3d 256 > gpurun_out/tune_v$v.log; done
# (each binary takes ~2 s per problem; the cks column must agree between variants).
# The PDL experiment is a library build: KSFD_NVCC_EXTRA=-DKSFD_PDL=1 python -m ksfd_b200.build --force,
# then KSFD_PDL=1 python -m pytest tests -m gpu -q && KSFD_PDL=0/1 python scripts/step_time.py
set -e
cd "$(dirname "$0")/.."
mkdir -p bin
for v in 0 1 2 4; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DKSFD_MARCH_VARIANT=$v \
       -I ksfd_b200/csrc scripts/tune_march.cu -o bin/tune_v$v &
done
wait
ls -la bin/tune_v*
