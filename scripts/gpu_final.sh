#!/bin/bash
# final round numbers (one B200): GPU tests, default bench line, reference arm, ncu launch list of the quick bench
# command, ncu --set full of the hot kernels (each only after its own command has exited 0 without ncu)
#   gpurun --timeout 1500 -- bash scripts/gpu_final.sh r02
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_gputests.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/${TAG}_gputests.log
timeout 600 python bench.py > gpurun_out/${TAG}_final_bench_line.json 2> gpurun_out/${TAG}_final_bench.err
echo "bench rc $?"; cut -c1-300 gpurun_out/${TAG}_final_bench_line.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_final_bench_reference_arm.json 2> gpurun_out/${TAG}_final_ref.err
echo "ref rc $?"
python bench.py --steps 2 --warmup 1 --quick --no-cpu > gpurun_out/${TAG}_quick.json 2> gpurun_out/${TAG}_quick.err && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_final_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --quick --no-cpu > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc $?"; wc -l gpurun_out/${TAG}_final_bench_launches.csv
for w in 2d1024 3d256; do
  python scripts/profile_target.py $w > /dev/null 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_tma_march|k_march" --launch-skip 8 --launch-count 4 \
      -o gpurun_out/${TAG}_prof_$w python scripts/profile_target.py $w > gpurun_out/${TAG}_ncu_$w.log 2>&1
  echo "ncu full $w rc $?"
done
