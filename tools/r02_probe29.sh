#!/bin/bash
# k_coeff occupancy variants (-DGM_COEFF_MINB=5 / 6: 96 / 80 registers with spills) against the default (4: 128 registers)
set -u
cd "$(dirname "$0")/.."
line() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=d['roofline']['kernel_ms_per_step']
print('$1 step %.3f' % d['ms_per_step'], {a: round(b, 3) for a, b in k.items() if a != 'launches' and a.startswith('k_')})"; }
for v in ${VARIANTS:-default minb5 minb6 default}; do
  if [ $v = default ]; then unset GEOSMIE_B200_LIB; else export GEOSMIE_B200_LIB=tools/variants/lib_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-lut --workloads su --steps 20 2>/dev/null | line su_$v
  timeout 300 python bench.py --no-cpu-baseline --no-lut --workloads ss --steps 3 2>/dev/null | line ss_$v
done
