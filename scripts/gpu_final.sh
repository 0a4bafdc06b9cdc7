#!/bin/bash
# final round numbers: default bench line, reference arm, ncu launch list of the quick bench command
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err
echo "bench rc $?"; cut -c1-400 gpurun_out/bench_r1_final2.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_final2_ref.json 2> gpurun_out/bench_r1_final2_ref.err
echo "ref rc $?"; cut -c1-400 gpurun_out/bench_r1_final2_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1_final2.csv python bench.py --steps 2 --warmup 1 --quick --no-cpu > gpurun_out/ncu_list_final2.log 2>&1
echo "ncu rc $?"; wc -l gpurun_out/launches_r1_final2.csv
