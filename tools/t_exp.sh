#!/bin/bash
# Dev harness (GPU box): parity then timing of the strip kernels.
for v in 5; do BRV_TC_VARIANT=$v timeout 200 python tools/fold_check.py fwd > gpurun_out/t${v}_fwd.log 2>&1; echo "variant $v rc $? ok $(grep -c 'ok ' gpurun_out/t${v}_fwd.log) bad $(grep -c BAD gpurun_out/t${v}_fwd.log)"; grep BAD gpurun_out/t${v}_fwd.log | head -3; done
timeout 600 python -m pytest tests/test_gpu_strip_kernels.py -q -x -k "forward or gradient or conv" 2>&1 | tail -3
FOLD_CHECK_VARIANTS=0,4 timeout 200 python tools/fold_check.py bench 2>&1 | grep -E "time" | grep cfg5 | cut -c1-100
