#!/bin/bash
# Dev harness (GPU box): parity then timing after code-size changes.
for v in 0 6; do BRV_TC_VARIANT=$v timeout 200 python tools/fold_check.py inv > gpurun_out/t${v}_inv.log 2>&1; echo "inv variant $v rc $? ok $(grep -c 'ok ' gpurun_out/t${v}_inv.log) bad $(grep -c BAD gpurun_out/t${v}_inv.log)"; done
timeout 600 python -m pytest tests/test_gpu_strip_kernels.py -q -x -k "inverse" 2>&1 | tail -2
FOLD_CHECK_VARIANTS=0,7 timeout 200 python tools/fold_check.py bench 2>&1 | grep -E "time" | cut -c1-100
