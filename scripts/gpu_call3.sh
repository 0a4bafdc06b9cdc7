#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/gpu_tests3.log 2>&1
echo "tests rc $?"; tail -15 gpurun_out/gpu_tests3.log
timeout 200 python scripts/profile_steps.py > gpurun_out/warm_profile3.log 2>&1
echo "profile rc $?"; head -12 gpurun_out/warm_profile3.log
timeout 300 python scripts/run_options84.py 400 > gpurun_out/opt84_auto400_scaled.log 2>&1
echo "auto400 rc $?"; tail -4 gpurun_out/opt84_auto400_scaled.log
