#!/bin/bash
# Dev harness (GPU box): forward kernel timing with pipeline roles disabled.
SK="-DBRV_T_NO_BUILD -DBRV_T_NO_EPI -DBRV_T_NO_TMA"
for flag in "$SK" "$SK -DBRV_T_ONE_PRODUCT" "$SK -DBRV_T_NO_MMA" "-DBRV_T_NO_MMA" "-DBRV_T_ONE_PRODUCT"; do
  NVCC_EXTRA="$flag" python __graft_entry__.py --force > /dev/null 2>&1
  echo "== flags: $flag"
  FOLD_CHECK_VARIANTS=0 python tools/fold_check.py bench 2>&1 | grep -E "cfg2|cfg5" | cut -c1-60
done
