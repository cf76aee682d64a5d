#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
line() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=d['roofline']['kernel_ms_per_step']
print('$1 step %.3f e2e %.3f k_coeff %.3f k_small %.3f k_gram %.3f sum+eval %.3f fin %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], k['k_coeff'], k.get('k_small',0), k['k_gram'], k['k_gram_sum_eval'], k['k_finalize']))"; }
B="--no-cpu-baseline --no-lut --workloads su --steps 20"
timeout 200 python bench.py $B 2>gpurun_out/p4_new.err | tee gpurun_out/p4_new.json | line new
GEOSMIE_NO_SMALL=1 timeout 200 python bench.py $B 2>gpurun_out/p4_old.err | tee gpurun_out/p4_old.json | line r01path
for sp in su bc ss; do timeout 300 python tools/profile_lut.py $sp > gpurun_out/p4_lut_$sp.txt 2>&1; head -2 gpurun_out/p4_lut_$sp.txt; done
sed -n 3,40p gpurun_out/p4_lut_ss.txt
echo "== ncu slow vs fast build of k_coeff (round-1 path, full optics_SU launch)"
GEOSMIE_NO_SMALL=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_coeff" -s 4 -c 1 -f -o gpurun_out/r02c_coeff_default_build_su_2196cells \
   python tools/one_step.py su 2 > gpurun_out/p4_ncu1.log 2>&1; grep -c "Profiling" gpurun_out/p4_ncu1.log
GEOSMIE_NO_SMALL=1 GEOSMIE_B200_LIB=tools/variants/lib_tpc4.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_coeff" -s 4 -c 1 -f -o gpurun_out/r02c_coeff_tpc4_build_su_2196cells \
   python tools/one_step.py su 2 > gpurun_out/p4_ncu2.log 2>&1; grep -c "Profiling" gpurun_out/p4_ncu2.log
