#!/usr/bin/env python
"""Per-function executed-instruction and stall-sample shares of one kernel (see ncu_lines.py).

    python scripts/ncu_funcs.py <report.ncu-rep> <kernel-regex> <lib.so> <rows>
"""
import bisect
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ncu_lines  # noqa: E402


def funcs_of(path):
    out = []
    for i, l in enumerate(open(path), 1):
        m2 = re.search(r"\b(\w+)\s*\(", l)
        if re.match(r"(SSDE_HD|__device__ __forceinline__|__global__)", l) and m2:
            out.append((i, m2.group(1)))
    return out


def main():
    rep, kre, lib, rows = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    inst, stall = ncu_lines.collect(rep, kre, lib)
    csrc = os.path.join(os.path.dirname(HERE), "smoothsde_b200", "csrc")
    tabs = {f: funcs_of(os.path.join(csrc, f)) for f in os.listdir(csrc) if f.endswith(".cuh")}
    agg, sagg = {}, {}
    ts = sum(stall.values())
    for key, v in inst.items():
        if key is None:
            name = "(none)"
        else:
            f, l = key
            if f in tabs and tabs[f]:
                starts = [s for s, _ in tabs[f]]
                k = bisect.bisect_right(starts, l) - 1
                name = f + ":" + (tabs[f][k][1] if k >= 0 else "?")
            else:
                name = f
        agg[name] = agg.get(name, 0) + v * 32 / rows
        sagg[name] = sagg.get(name, 0) + stall[key] / ts * 100
    print(f"total {sum(agg.values()):.0f} thread-instructions per row")
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:24]:
        print(f"{n:48s} {v:7.1f} instr/row   stall samples {sagg[n]:5.1f}%")


if __name__ == "__main__":
    main()
