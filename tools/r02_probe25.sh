#!/bin/bash
# device-buffer cache: GPU tests, then optics_SS builds with and without the cache in one call
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== cache on";  python tools/diag_lut_outliers.py ss 2>&1 | grep "^run"
echo "== cache off"; GEOSMIE_NO_ALLOC_CACHE=1 python tools/diag_lut_outliers.py ss 2>&1 | grep "^run\|close\|__init__"
echo "== cache on";  python tools/diag_lut_outliers.py ss 2>&1 | grep "^run\|close\|__init__"
timeout 300 python bench.py --no-cpu-baseline --no-lut --workloads su --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('su step %.3f e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
