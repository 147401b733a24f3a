"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel family.

    python scripts/launch_shares.py gpurun_out/launches.csv [skip_launches] > profiles/rNN_launch_shares.json

Per-launch times under ncu are cold-cache and serialised: the SHARES are what carries over to the un-profiled step.
"""
import csv
import json
import re
import sys
from collections import defaultdict


def family(name):
    name = re.sub(r"^void ", "", name)
    if name.startswith("at::") or "at::native" in name:
        m = re.search(r"at::native::([A-Za-z0-9_]+)", name)
        return "torch:" + (m.group(1) if m else re.sub(r"[<(].*$", "", name))[:60]
    name = name.replace("vsw::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    m = re.match(r"([A-Za-z0-9_]+)(<[^>]*>)?", name)
    base = m.group(1)
    if base == "tc_gemm_kernel" and m.group(2):
        return base + m.group(2).replace(" ", "")
    return base


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1000.0 if unit in ("ns", "nsecond") else v * 1000.0 if unit in ("ms", "msecond") else v
        rows.append((int(r["ID"]), r["Kernel Name"], v_us))
    rows = [r for r in rows if r[0] >= skip]
    tot = sum(r[2] for r in rows)
    fam = defaultdict(lambda: [0, 0.0])
    for _, n, v in rows:
        k = family(n)
        fam[k][0] += 1
        fam[k][1] += v
    out = {"launches": len(rows), "total_us": round(tot, 1),
           "families": {k: {"launches": c, "us": round(v, 1), "share": round(v / tot, 4)}
                        for k, (c, v) in sorted(fam.items(), key=lambda kv: -kv[1][1])}}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
