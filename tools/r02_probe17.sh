#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
line() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=d['roofline']['kernel_ms_per_step']
print('$1 step %.3f e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {a: round(b, 3) for a, b in k.items() if a != 'launches' and a.startswith('k_')})"; }
B="--no-cpu-baseline --no-lut --workloads su --steps 20"
timeout 300 python bench.py $B 2>/dev/null | line su_nodes
GEOSMIE_EVAL_DIRECT=1 timeout 300 python bench.py $B 2>/dev/null | line su_direct
timeout 300 python bench.py --no-cpu-baseline --no-lut --workloads ss --steps 3 2>/dev/null | line ss_nodes
