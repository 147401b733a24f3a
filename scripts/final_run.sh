#!/bin/bash
# single-GPU evidence run of a round: GPU tests, smoke, the contract bench line and its variants, ncu launch list
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/${R}_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${R}_final_smoke.log 2>&1
python bench.py 2>$O/${R}_bench_default.err | tail -1 > $O/${R}_final_bench_default.json
python bench.py --dtype fp16 --no-cpu-baseline 2>/dev/null | tail -1 > $O/${R}_final_bench_fp16.json
python bench.py --model violet --no-cpu-baseline 2>/dev/null | tail -1 > $O/${R}_final_bench_violet_b32.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/${R}_final_bench_reference_arm.json
python bench.py --model swin_l_384 --batch 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>$O/${R}_swin_l.err | tail -1 > $O/${R}_final_bench_swin_l_384_b2.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 6300 -c 1150 --csv --log-file $O/${R}_final_launches_bench_b32.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/${R}_ncu_launches.log 2>&1
tail -3 $O/${R}_final_pytest_gpu.log; cat $O/${R}_final_smoke.log | tail -1
for f in default fp16 violet_b32 reference_arm swin_l_384_b2; do python - <<PY
import json
try:
    d = json.load(open("$O/${R}_final_bench_$f.json"))
    print("$f", round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms", (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), d.get("clocks"))
except Exception as e:
    print("$f", "FAILED", e)
PY
done
