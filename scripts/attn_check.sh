#!/bin/bash
set -u
export VSW_ATTN_TC2=3
for args in "8 56 56 8 7 7 0 0 0 4 32 bf16 3" "8 14 14 8 7 7 0 3 3 16 32 bf16 3" "4 12 12 4 6 6 2 3 3 4 4 bf16 1" "8 28 28 8 7 7 0 3 3 8 8 fp16 1"; do
  echo "=== $args"
  DBG_BWD=1 timeout 120 python scripts/dbg_attn.py $args 2>&1 | grep -v "^Traceback\|^  File" | tail -5 | cut -c1-300
done
