#!/bin/bash
# one GPU call: spectral-preconditioner test, timing table, options84 with each preconditioner
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k spectral -x > gpurun_out/fft_test.log 2>&1
echo "test rc $?"; tail -15 gpurun_out/fft_test.log
timeout 300 python scripts/debug_fftpc.py > gpurun_out/fft_debug.log 2>&1
echo "debug rc $?"; cat gpurun_out/fft_debug.log | tail -30
timeout 400 python scripts/run_options84.py 100 > gpurun_out/opt84_auto.log 2>&1
echo "auto rc $?"; tail -4 gpurun_out/opt84_auto.log
timeout 400 python scripts/run_options84.py 100 pbjacobi > gpurun_out/opt84_pbj.log 2>&1
echo "pbj rc $?"; tail -4 gpurun_out/opt84_pbj.log
