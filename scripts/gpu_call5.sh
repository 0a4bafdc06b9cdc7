#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/gpu_multi5.log 2>&1
echo "multi tests rc $?"; tail -5 gpurun_out/gpu_multi5.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 3 --quick --no-cpu > gpurun_out/bench_n2_r1q.json 2> gpurun_out/bench_n2_r1q.err
echo "bench n2 rc $?"; cat gpurun_out/bench_n2_r1q.json | cut -c1-600
