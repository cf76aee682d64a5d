#!/bin/bash
# k_coeff instantiation choice: forced short (128 registers) / forced long (168 registers) / automatic, GPU tests with the automatic choice
set -u
cd "$(dirname "$0")/.."
line() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=d['roofline']['kernel_ms_per_step']
print('$1 step %.3f' % d['ms_per_step'], {a: round(b, 3) for a, b in k.items() if a != 'launches' and a.startswith('k_')})"; }
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
for v in 0 1 auto; do
  if [ $v = auto ]; then unset GEOSMIE_COEFF_LONG; else export GEOSMIE_COEFF_LONG=$v; fi
  timeout 300 python bench.py --no-cpu-baseline --no-lut --workloads su --steps 20 2>/dev/null | line su_long$v
  timeout 300 python bench.py --no-cpu-baseline --no-lut --workloads ss --steps 3 2>/dev/null | line ss_long$v
done
