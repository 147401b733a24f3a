#!/bin/bash
# timing of the attention kernels at the four Swin-B stage shapes (+ correctness line), twice
set -u
for rep in 1 2; do
for args in "8 56 56 8 7 7 0 0 0 4 32 bf16 10" "8 28 28 8 7 7 0 3 3 8 32 bf16 10" "8 14 14 8 7 7 0 3 3 16 32 bf16 10" "8 7 7 8 7 7 0 0 0 32 32 bf16 10"; do
  DBG_BWD=${DBG_BWD:-0} timeout 120 python scripts/dbg_attn.py $args 2>&1 | tail -${TAILN:-1}
done; done
