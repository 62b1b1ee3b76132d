#!/usr/bin/env python
"""Attribute executed warp instructions and stall samples of one kernel to source lines.

    python scripts/ncu_lines.py <report.ncu-rep | NAME.source.csv> <kernel-regex> <lib.so> [rows]

ncu's SASS page (per-instruction counters) is joined, by instruction order, with nvdisasm's
line-info disassembly of the same cubin.  `rows` scales the counts to instructions per row.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def _collect(rep, kre, lib):
    if rep.endswith(".csv"):      # the SASS page dumped on the GPU box (scripts/ncu_capture.sh: NAME.source.csv)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                              "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
    lines = raw.splitlines()
    kstart = next((i for i, l in enumerate(lines) if l.startswith('"Kernel Name"') and re.search(kre, l)), 0)
    start = next(i for i, l in enumerate(lines) if i > kstart - 1 and l.startswith('"Address"'))
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
    kname = lines[start - 1]
    rd = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
    hdr = rd[0]
    ix = {h: i for i, h in enumerate(hdr)}
    sass = rd[1:]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    dis = []
    for f in sorted(os.listdir(tmp)):
        if f.endswith(".cubin"):
            dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines()
    funcs = {}
    cur, curline = None, None
    for l in dis:
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            curline = None
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m and cur:
            curline = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            funcs[cur].append(curline)
    cands = [f for f, v in funcs.items() if len(v) == len(sass) and re.search(kre, f)]
    if not cands:
        cands = [f for f, v in funcs.items() if len(v) == len(sass)]
    if not cands:
        raise SystemExit(f"no function with {len(sass)} instructions in {lib} (library rebuilt since the capture?); "
                         f"candidates: { {f: len(v) for f, v in funcs.items() if re.search(kre, f)} }")
    lineinfo = funcs[cands[0]]
    inst, stall, longsb = defaultdict(float), defaultdict(float), defaultdict(float)
    for k, r in enumerate(sass):
        key = lineinfo[k]
        inst[key] += float(r[ix["Instructions Executed"]] or 0)
        stall[key] += float(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
        longsb[key] += float(r[ix["stall_long_sb"]] or 0)
    return inst, stall, longsb, kname


def collect(rep, kre, lib):
    """-> (instructions executed, stall samples) keyed by (file, line)."""
    r = _collect(rep, kre, lib)
    return r[0], r[1]


def main():
    rep, kre, lib = sys.argv[1:4]
    rows = float(sys.argv[4]) if len(sys.argv) > 4 else None
    inst, stall, longsb, kname = _collect(rep, kre, lib)
    tot_i, tot_s = sum(inst.values()), sum(stall.values())
    print(kname[:120])
    print(f"total warp instructions {tot_i:.0f}" + (f" = {tot_i * 32 / rows:.0f} thread-instr/row" if rows else "")
          + f", stall samples {tot_s:.0f}")
    print("top lines by stall samples:")
    for key, v in sorted(stall.items(), key=lambda kv: -kv[1])[:25]:
        print(f"  {str(key):40s} stalls {100 * v / tot_s:5.1f}% (long_sb {100 * longsb[key] / tot_s:5.1f}%)  instr {100 * inst[key] / tot_i:5.1f}%")
    print("top lines by instructions:")
    for key, v in sorted(inst.items(), key=lambda kv: -kv[1])[:25]:
        print(f"  {str(key):40s} instr {100 * v / tot_i:5.1f}%" + (f" = {v * 32 / rows:6.1f}/row" if rows else ""))


if __name__ == "__main__":
    main()
