#!/bin/bash
mkdir -p gpurun_out
KSFD_GM_SPECULATE=0 timeout 120 python scripts/step_time.py > gpurun_out/spec_ab.log 2>&1
KSFD_GM_SPECULATE=1 timeout 120 python scripts/step_time.py >> gpurun_out/spec_ab.log 2>&1
KSFD_GM_SPECULATE=0 timeout 120 python scripts/step_time.py >> gpurun_out/spec_ab.log 2>&1
KSFD_GM_SPECULATE=1 timeout 120 python scripts/step_time.py >> gpurun_out/spec_ab.log 2>&1
cat gpurun_out/spec_ab.log
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/gpu_tests4.log 2>&1
echo "tests rc $?"; tail -5 gpurun_out/gpu_tests4.log
