#!/bin/bash
# ncu --set full capture summarised ON the GPU box: the .ncu-rep stays in /tmp (with --import-source
# on it is tens of MB per kernel, and gpurun_out/ is limited to 64 MiB); only the markdown summary
# (scripts/ncu_summary.py) and the raw per-launch CSV come back.
# usage: [NCU_SKIP=n] scripts/ncu_capture.sh NAME 'KERNEL_REGEX' COUNT command...   (NCU_SKIP: matching launches to skip first)
name=$1; regex=$2; count=$3; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$regex" --launch-skip "${NCU_SKIP:-0}" -c "$count" -f -o /tmp/$name "$@" > gpurun_out/$name.stdout 2> gpurun_out/$name.stderr
python scripts/ncu_summary.py /tmp/$name.ncu-rep > gpurun_out/$name.md 2>> gpurun_out/$name.stderr
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>> gpurun_out/$name.stderr
ncu -i /tmp/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/$name.source.csv 2>> gpurun_out/$name.stderr   # per-instruction counters: scripts/ncu_lines.py
ls -la /tmp/$name.ncu-rep >> gpurun_out/$name.stderr
