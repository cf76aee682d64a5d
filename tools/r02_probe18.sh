#!/bin/bash
# k_gram with symmetric halves / ragged-tile skipping per team size (GM_GRAM_SYMMETRIC bit mask) against the default build
set -u
cd "$(dirname "$0")/.."
line() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=d['roofline']['kernel_ms_per_step']
print('$1 step %.3f' % d['ms_per_step'], {a: round(b, 3) for a, b in k.items() if a != 'launches' and a.startswith('k_')})"; }
B="--no-cpu-baseline --no-lut --workloads su --steps 20"
timeout 300 python bench.py $B 2>/dev/null | line default
for m in 1 2 4 8 12; do
  GEOSMIE_B200_LIB=tools/variants/libgm_sym$m.so timeout 300 python bench.py $B 2>/dev/null | line sym$m
done
timeout 300 python bench.py $B 2>/dev/null | line default_again
