#!/bin/bash
# Dev harness (GPU box): parity then timing of the inverse strip kernel.
for v in 0; do BRV_TC_VARIANT=$v timeout 200 python tools/fold_check.py inv > gpurun_out/t${v}_inv.log 2>&1; echo "inv variant $v rc $? ok $(grep -c 'ok ' gpurun_out/t${v}_inv.log) bad $(grep -c BAD gpurun_out/t${v}_inv.log)"; done
timeout 600 python -m pytest tests/test_gpu_strip_kernels.py tests/test_gpu_tensorcore.py -q -x -k "inverse or gradient or istft" 2>&1 | tail -2
FOLD_CHECK_VARIANTS=0 timeout 200 python tools/fold_check.py bench 2>&1 | grep -E "time" | cut -c1-100
