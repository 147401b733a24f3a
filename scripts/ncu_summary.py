"""Summarise an `ncu --set full` report (.ncu-rep) per kernel: durations, pipe utilisation, shared-memory wavefronts, DRAM bytes.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > x.csv ;  python scripts/ncu_summary.py x.csv "<command that made it>" > profiles/rNN_ncu_x.json
"""
import csv
import json
import sys
from collections import defaultdict

WANT = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed": "smem_lsu_wavefront_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_lsu_wavefronts",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__inst_executed.avg.per_cycle_active": "ipc_per_smsp",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc_sm",
    "launch__registers_per_thread": "registers",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "smsp__cycles_active.avg": "cycles_active",
}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kcol = col["Kernel Name"]
    out = defaultdict(lambda: defaultdict(list))
    for r in data:
        name = r[kcol].split("(")[0].replace("void ", "").replace("vsw::", "").replace("<unnamed>::", "")
        for m, key in WANT.items():
            if m in col:
                v = num(r[col[m]])
                if v is not None:
                    u = units[col[m]]
                    if key == "duration_ns" and u in ("us", "usecond"):
                        v *= 1e3
                    if key == "duration_ns" and u in ("ms", "msecond"):
                        v *= 1e6
                    if key.endswith("_bytes"):
                        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
                    out[name][key].append(v)
    res = {"command": sys.argv[2] if len(sys.argv) > 2 else None, "kernels": {}}
    for name, d in out.items():
        n = len(d["duration_ns"])
        e = {"launches": n, "sum_ms": sum(d["duration_ns"]) / 1e6}
        for k, v in d.items():
            if k == "duration_ns":
                continue
            e["avg_" + k] = sum(v) / len(v)
        if "dram_read_bytes" in d:
            e["avg_dram_bytes_per_launch"] = (sum(d["dram_read_bytes"]) + sum(d["dram_write_bytes"])) / n
        res["kernels"][name] = e
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
