#!/bin/bash
# attention-focused GPU check: kernel tests, model tests, micro-bench
set -u
mkdir -p gpurun_out
echo "== pytest attention =="
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 300 -k "attention" > gpurun_out/pytest_attn.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_attn.log
tail -30 gpurun_out/pytest_attn.log
echo "== kbench attn =="
timeout 300 python scripts/kbench.py attn > gpurun_out/kbench_attn.txt 2>&1
tail -9 gpurun_out/kbench_attn.txt
if [ "${1:-}" = "full" ]; then
echo "== pytest all =="
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
fi
