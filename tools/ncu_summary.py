#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the per-launch JSON + markdown table kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_xxx.json

Reads the report with `ncu -i <rep> --page raw --csv` (works on the CPU box) and keeps, per launch: kernel name,
grid, duration, DRAM bytes read / written, DRAM %, tensor-pipe %, XU (MUFU) %, issue-active %, SM %, L1TEX %,
registers per thread.
"""
import csv
import io
import json
import re
import subprocess
import sys

COLS = {
    "us": ("gpu__time_duration.sum", 1e-3),
    "dram_read_mb": ("dram__bytes_read.sum", None),
    "dram_write_mb": ("dram__bytes_write.sum", None),
    "dram_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "tensor_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    "xu_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
    "issue_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    "sm_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "l1tex_pct": ("l1tex__throughput.avg.pct_of_peak_sustained_active", 1.0),
    "regs": ("launch__registers_per_thread", 1.0),
}
UNIT_TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def short_name(full: str) -> str:
    name = full.split("(")[0].strip()
    return re.sub(r"^void\s+", "", name)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        grid = r[idx["Grid Size"]] if "Grid Size" in idx else r[idx["launch__grid_size"]]
        nums = [int(x) for x in re.findall(r"\d+", grid)]
        g = 1
        for v in nums:
            g *= v
        rec = {"kernel": short_name(r[idx["Kernel Name"]]), "grid": g}
        for key, (metric, scale) in COLS.items():
            if metric not in idx:
                continue
            val = float(r[idx[metric]].replace(",", ""))
            if key == "us":
                val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[idx[metric]], 1e-3)
            elif scale is None:
                val *= UNIT_TO_MB.get(units[idx[metric]], 1e-6)
            else:
                val *= scale
            rec[key] = round(val, 2)
        res.append(rec)
    json.dump(res, open(out, "w"), indent=1)
    keys = ["kernel", "grid", "us", "dram_read_mb", "dram_write_mb", "dram_pct", "tensor_pct", "xu_pct", "issue_pct", "sm_pct", "regs"]
    print("| " + " | ".join(keys) + " |")
    print("|" + "---|" * len(keys))
    for rec in res:
        print("| " + " | ".join(str(rec.get(k, "")) for k in keys) + " |")
    print(f"\nsum of durations: {sum(r['us'] for r in res) / 1000:.3f} ms over {len(res)} launches")


if __name__ == "__main__":
    main()
