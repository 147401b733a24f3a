#!/bin/bash
set -u
export VSW_ATTN_TC2=1
for args in "8 7 7 8 7 7 0 0 0 2 1 bf16 0" "8 14 14 8 7 7 0 3 3 2 1 bf16 0" "4 12 12 4 6 6 0 3 3 4 2 bf16 0" "8 56 56 8 7 7 0 0 0 4 32 bf16 5" "8 14 14 8 7 7 0 3 3 16 32 bf16 5"; do
  echo "=== $args"
  DBG_BWD=1 timeout 120 python scripts/dbg_attn.py $args 2>&1 | grep -v "^Traceback\|^  File\|^    " | tail -6
done
