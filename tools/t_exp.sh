#!/bin/bash
# Dev harness (GPU box): back-off cap of the relaxed mbarrier waits.
for flag in "-DBRV_WAIT_NAP_MAX=64u" "-DBRV_WAIT_NAP_MAX=256u" "-DBRV_WAIT_NAP_MAX=1024u" "-DBRV_WAIT_NAP_MAX=4096u"; do
  NVCC_EXTRA="$flag" python __graft_entry__.py --force > /dev/null 2>&1
  echo "== flags: $flag"
  FOLD_CHECK_VARIANTS=0,5,6 python tools/fold_check.py bench 2>&1 | grep -E "time" | grep -v "cfg1" | cut -c1-100
done
