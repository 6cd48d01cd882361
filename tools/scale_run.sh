#!/bin/bash
# Scaling run on ONE 8-GPU box: the bench at N = 1, 2, 4, 8 launched exactly like the driver does.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/scale_run.sh'
OUT=gpurun_out
mkdir -p $OUT
for w in cfg2 cfg5 cfg3; do
  for n in 1 2 4 8; do
    if [ $n = 1 ]; then
      timeout 200 python bench.py --gpus 1 --workload $w --no-cpu-baseline > $OUT/scale_${w}_n1.json 2> $OUT/scale_${w}_n1.err
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port $((29500 + n)) bench.py --gpus $n --workload $w > $OUT/scale_${w}_n$n.json 2> $OUT/scale_${w}_n$n.err
    fi
    python - <<PY
import json
try:
    d = json.loads(open('$OUT/scale_${w}_n$n.json').read().strip().splitlines()[-1])
    print('$w', 'n=$n', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print('$w n=$n failed', e)
PY
  done
done
