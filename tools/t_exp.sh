#!/bin/bash
# Dev harness (GPU box): parity then timing of the forward strip kernel flavours.
for v in 5 8; do BRV_TC_VARIANT=$v timeout 200 python tools/fold_check.py fwd > gpurun_out/t${v}_fwd.log 2>&1; echo "fwd variant $v rc $? ok $(grep -c 'ok ' gpurun_out/t${v}_fwd.log) bad $(grep -c BAD gpurun_out/t${v}_fwd.log)"; grep BAD gpurun_out/t${v}_fwd.log | head -3; done
FOLD_CHECK_VARIANTS=4,5,8 timeout 200 python tools/fold_check.py bench 2>&1 | grep -E "time" | cut -c1-60
