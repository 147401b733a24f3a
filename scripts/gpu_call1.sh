#!/bin/bash
# round-2 call 1: TMEM micro-benchmarks, full GPU test-suite (incl. the new full-size parity tests), phase counters of the
# round-1 attention kernels, and the round-1 bench line as the baseline of this round.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== ubench tmem =="
timeout 120 scripts/ubench/tmem > gpurun_out/ubench_tmem.txt 2>&1; echo "exit $?" >> gpurun_out/ubench_tmem.txt
cat gpurun_out/ubench_tmem.txt
echo "== pytest -m gpu =="
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
echo "== attention phase counters (round-1 kernels) =="
VSW_ATTN_DEBUG=1 VSW_ATTN_DEBUG_DUMP=1 timeout 120 python scripts/prof_attn.py 8 > gpurun_out/attn_debug.txt 2>&1
tail -12 gpurun_out/attn_debug.txt
echo "== kbench attn =="
timeout 300 python scripts/kbench.py attn > gpurun_out/kbench_attn.txt 2>&1
tail -12 gpurun_out/kbench_attn.txt
echo "== bench =="
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_start.json 2> gpurun_out/bench_r02_start.err
tail -c 3000 gpurun_out/bench_r02_start.json
