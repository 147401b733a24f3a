#!/bin/bash
set -u
echo "== old kernels with elect_one =="
VSW_ATTN_TC2=0 timeout 300 python scripts/kbench.py attn 2>&1 | tail -7 | cut -c1-260
echo "== pytest kernels (old attention path + gemm) =="
VSW_ATTN_TC2=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 300 -k "attention or linear" 2>&1 | tail -4
echo "== kbench linear =="
timeout 300 python scripts/kbench.py linear 2>&1 | tail -12 | cut -c1-400
