#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
line() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=d['roofline']['kernel_ms_per_step']
print('$1 step %.2f e2e %.2f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {a: round(b, 2) for a, b in k.items() if a != 'launches' and a.startswith('k_')})"; }
B="--no-cpu-baseline --no-lut --workloads ss --steps 3"
timeout 300 python bench.py $B 2>/dev/null | line ss_default
GEOSMIE_B200_LIB=tools/variants/lib_nostore.so timeout 300 python bench.py $B 2>/dev/null | line ss_nostore
GEOSMIE_B200_LIB=tools/variants/lib_stcs.so timeout 300 python bench.py $B 2>/dev/null | line ss_stcs
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_coeff|k_contract" -s 2 -c 2 -f -o gpurun_out/r02d_coeff_contract_ss_bin4 \
   python tools/prof_case.py ss 4 65 > gpurun_out/p8_ncu.log 2>&1; tail -2 gpurun_out/p8_ncu.log | cut -c1-300
