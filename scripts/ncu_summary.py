#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small markdown table.

    python scripts/ncu_summary.py gpurun_out/r2_full.ncu-rep > profiles/r01_xxx.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    print(f"# ncu summary of `{rep}`\n")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        print(f"## {name[:100]}\n")
        print("| metric | value |\n|---|---|")
        for k, lab in KEYS:
            if k in idx:
                print(f"| {lab} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |")
        if "dram__bytes_read.sum" in idx:
            def tobytes(k):
                v, u = float(r[idx[k]]), units[idx[k]].lower()
                return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            tot = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
            t = float(r[idx["gpu__time_duration.sum"]])
            tu = units[idx["gpu__time_duration.sum"]]
            ts = t * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(tu, 1e-6)
            print(f"| dram traffic (read+write) | {tot / 1e9:.4f} GB -> {tot / ts / 1e9:.0f} GB/s under ncu |")
        top = sorted(((float(r[idx[h]]), h.split("stalled_")[1].split("_per_")[0]) for h in stall), reverse=True)[:5]
        print("| top stalls (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in top) + " |")
        print()


if __name__ == "__main__":
    main()
