#!/bin/bash
# Dev harness (GPU box): builder register-ring depth of the strip inverse kernel (timing only).
for flag in "-DBRV_T_NB=1" "-DBRV_T_NB=2"; do
  NVCC_EXTRA="$flag" python __graft_entry__.py --force > /dev/null 2>&1
  echo "== flags: $flag"
  FOLD_CHECK_VARIANTS=0 python tools/fold_check.py bench 2>&1 | grep -E "time" | cut -c1-100
done
