#!/bin/bash
# Round-end style scaling run on ONE box with 8 GPUs: multi-GPU parity check on 8 ranks, then the weak-scaling
# bench at 1/2/4/8 GPUs (quick), then the full 8-GPU line (parity + strong-scaling records).
#   gpurun --gpus 8 -- bash scripts/gpu_scale.sh
set -u
OUT=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $T --nproc-per-node 8 --master-port 29801 scripts/multi_gpu_check.py > $OUT/mgc8.log 2>&1
echo "multi_gpu_check(8) exit $?"; grep -E "MULTI_GPU_CHECK|FAIL" $OUT/mgc8.log | head
python bench.py --steps 20 --warmup 3 --quick --no-cpu > $OUT/scale_n1.json 2> $OUT/scale_n1.err; echo "n1 $?"
for n in 2 4 8; do
  timeout 400 $T --nproc-per-node $n --master-port $((29810 + n)) bench.py --gpus $n --steps 20 --warmup 3 --quick --no-cpu \
      > $OUT/scale_n$n.json 2> $OUT/scale_n$n.err; echo "n$n $?"
done
timeout 600 $T --nproc-per-node 8 --master-port 29830 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu \
    > $OUT/scale_n8_full.json 2> $OUT/scale_n8_full.err; echo "n8 full $?"
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        t = open('gpurun_out/scale_n%d.json' % n).read(); j = json.loads(t[t.index('{'):])
        print(n, 'ms/step %.3f value %.1f launches/step %.0f parity %s' % (
            j['ms_per_step'], j['value'], j['gpu_launches'] / j['steps'], (j.get('parity') or {}).get('ok')))
    except Exception as e:
        print(n, 'ERR', e)
PY
