#!/bin/bash
# Dev harness (GPU box): parity then timing after code-size changes.
for v in 0; do BRV_TC_VARIANT=$v timeout 200 python tools/fold_check.py inv > gpurun_out/t${v}_inv.log 2>&1; echo "inv variant $v rc $? ok $(grep -c 'ok ' gpurun_out/t${v}_inv.log) bad $(grep -c BAD gpurun_out/t${v}_inv.log)"; done
for v in 0 5; do BRV_TC_VARIANT=$v timeout 200 python tools/fold_check.py fwd > gpurun_out/t${v}_fwd.log 2>&1; echo "fwd variant $v rc $? ok $(grep -c 'ok ' gpurun_out/t${v}_fwd.log) bad $(grep -c BAD gpurun_out/t${v}_fwd.log)"; done
FOLD_CHECK_VARIANTS=0,4 timeout 200 python tools/fold_check.py bench 2>&1 | grep -E "time" | cut -c1-100
