#!/bin/bash
# Dev harness (GPU box): waits with a try_wait suspend-time hint vs try_wait + nanosleep.
for flag in "" "-DBRV_WAIT_HINT"; do
  NVCC_EXTRA="$flag" python __graft_entry__.py --force > /dev/null 2>&1
  echo "== flags: $flag"
  FOLD_CHECK_VARIANTS=0,4 python tools/fold_check.py bench 2>&1 | grep -E "time" | cut -c1-100
done
