#!/bin/bash
# Dev harness (GPU box): parity then timing of the transposed inverse kernel variants.
for v in 7; do BRV_TC_VARIANT=$v timeout 200 python tools/fold_check.py inv > gpurun_out/t${v}_inv.log 2>&1; echo "variant $v rc $? ok $(grep -c 'ok ' gpurun_out/t${v}_inv.log) bad $(grep -c BAD gpurun_out/t${v}_inv.log)"; grep BAD gpurun_out/t${v}_inv.log | head -5; done
timeout 600 python -m pytest tests/test_gpu_strip_kernels.py -q -x -k "inverse or gradient" 2>&1 | tail -3
FOLD_CHECK_VARIANTS=4,6,7 timeout 200 python tools/fold_check.py bench 2>&1 | grep -E "time" | grep "cfg5\|cfg1" | cut -c1-100
