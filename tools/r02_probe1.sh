#!/bin/bash
# round 2, call 1 (1 GPU): parity of the fused small-class kernel, step/kernel times new vs round-1 path vs staged k_coeff variants,
# ncu --set full of k_small and k_coeff.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,uuid,pci.bus_id,clocks.max.sm,ecc.mode.current,memory.used --format=csv > gpurun_out/p1_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
line() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=d['roofline']['kernel_ms_per_step']
print('$1 step %.3f e2e %.3f k_coeff %.3f k_small %.3f k_gram %.3f sum+eval %.3f fin %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], k['k_coeff'], k.get('k_small',0), k['k_gram'], k['k_gram_sum_eval'], k['k_finalize']))"; }
timeout 200 python bench.py --no-cpu-baseline --steps 20 2>gpurun_out/p1_new.err | tee gpurun_out/p1_new.json | line new
GEOSMIE_NO_SMALL=1 timeout 200 python bench.py --no-cpu-baseline --steps 20 2>gpurun_out/p1_old.err | tee gpurun_out/p1_old.json | line r01path
for v in tpc4 short nostore; do
  GEOSMIE_B200_LIB=tools/variants/lib_$v.so timeout 200 python bench.py --no-cpu-baseline --steps 20 2>gpurun_out/p1_$v.err | tee gpurun_out/p1_$v.json | line $v
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_small|k_coeff" -s 4 -c 2 -f -o gpurun_out/r02a_small_coeff_su_549cells \
   python tools/prof_case.py su 0 549 > gpurun_out/p1_ncu.log 2>&1
tail -3 gpurun_out/p1_ncu.log
GEOSMIE_NO_SMALL=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_coeff" -s 2 -c 1 -f -o gpurun_out/r02a_coeff_r01path_su_549cells \
   python tools/prof_case.py su 0 549 > gpurun_out/p1_ncu2.log 2>&1
tail -2 gpurun_out/p1_ncu2.log
timeout 200 python bench.py --no-cpu-baseline --steps 20 2>/dev/null | line new_again
