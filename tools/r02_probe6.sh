#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30
echo "== config 5 on 1 GPU"
timeout 600 python tools/config5_bands.py su bc ss 2>&1 | tail -4 | cut -c1-600
