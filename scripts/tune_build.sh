#!/bin/bash
# Kernel-tuning build: only the benchmark's kernels (CTCRW, d = 2, fp64), compiled in seconds into
# smoothsde_b200/lib/libsmoothsde_b200_<suffix>.so.
#   scripts/tune_build.sh <suffix> [extra nvcc flags, e.g. -DSSDE_KNT=64 -DSSDE_MINB=6]
#   SSDE_LIB_SUFFIX=_<suffix> python bench.py --no-cpu-baseline
set -e
cd "$(dirname "$0")/.."
SUF=$1; shift
SSDE_LIB_SUFFIX=_$SUF SSDE_NVCC_EXTRA="-DSSDE_MINIMAL $*" python -m smoothsde_b200.build --force
