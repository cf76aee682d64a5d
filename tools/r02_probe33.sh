#!/bin/bash
# prefetch depth of the long k_coeff build (4 / 6 / 8) on optics_SS
set -u
cd "$(dirname "$0")/.."
line() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=d['roofline']['kernel_ms_per_step']
print('$1 step %.3f' % d['ms_per_step'], {a: round(b, 3) for a, b in k.items() if a != 'launches' and a.startswith('k_')})"; }
for v in default pd6 pd8 default; do
  if [ $v = default ]; then unset GEOSMIE_B200_LIB; else export GEOSMIE_B200_LIB=tools/variants/lib_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-lut --workloads ss --steps 3 2>/dev/null | line ss_$v
done
