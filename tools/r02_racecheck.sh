#!/bin/bash
# compute-sanitizer synccheck and racecheck (shared-memory hazards) over the small table tests
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K="ragged or table_cells_against_oracle or device_normalisation or table_linearity"
for tool in synccheck racecheck; do
  ( time timeout 900 compute-sanitizer --tool $tool --print-limit 200 --log-file gpurun_out/r02_$tool.log \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" ) 2>&1 | tail -5
  tail -2 gpurun_out/r02_$tool.log
  grep -A2 "hazard\|Barrier error\|Divergent" gpurun_out/r02_$tool.log | grep " at " | sort | uniq -c | sort -rn | head -12
done
true
