#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
df -h /tmp /dev/shm | cat
echo "== /tmp, outputs deleted after each run"; python tools/diag_lut_outliers.py ss 2>&1 | tail -48
echo "== /dev/shm"; python tools/diag_lut_outliers.py ss /dev/shm 2>&1 | grep "^run"
