#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/profile_steps.py > gpurun_out/warm_profile.log 2>&1
echo "profile rc $?"; tail -40 gpurun_out/warm_profile.log
timeout 200 python scripts/run_options84.py 400 > gpurun_out/opt84_auto400.log 2>&1
echo "auto400 rc $?"; tail -5 gpurun_out/opt84_auto400.log
