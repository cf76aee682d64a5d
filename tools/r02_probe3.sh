#!/bin/bash
# round 2, call 3 (1 GPU): full GPU test-suite, the new bench line (SS headline + SU + lut_build_s + reference CPU arm), variants of the
# coefficient stores / k_small, ncu captures (slow build vs default, k_small), DRAM bytes per kernel.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
line() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=d['roofline']['kernel_ms_per_step']
print('$1 step %.3f e2e %.3f k_coeff %.3f k_small %.3f k_gram %.3f sum+eval %.3f fin %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], k['k_coeff'], k.get('k_small',0), k['k_gram'], k['k_gram_sum_eval'], k['k_finalize']))"; }
B="--no-cpu-baseline --no-lut --workloads su --steps 20"
timeout 200 python bench.py $B 2>gpurun_out/p3_new.err | tee gpurun_out/p3_new.json | line new
GEOSMIE_NO_SMALL=1 timeout 200 python bench.py $B 2>gpurun_out/p3_old.err | tee gpurun_out/p3_old.json | line r01path
for v in stcs smallshort minb2 tpc4; do
  GEOSMIE_B200_LIB=tools/variants/lib_$v.so timeout 200 python bench.py $B 2>gpurun_out/p3_$v.err | tee gpurun_out/p3_$v.json | line $v
done
GEOSMIE_NO_SMALL=1 GEOSMIE_B200_LIB=tools/variants/lib_stcs.so timeout 200 python bench.py $B 2>/dev/null | line stcs_r01path
GEOSMIE_NO_SMALL=1 GEOSMIE_B200_LIB=tools/variants/lib_tpc4.so timeout 200 python bench.py $B 2>/dev/null | line tpc4_r01path
echo "== full default bench"
( time timeout 900 python bench.py > gpurun_out/p3_bench_full.json 2> gpurun_out/p3_bench_full.err ) 2>&1 | grep real
tail -c 1500 gpurun_out/p3_bench_full.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/p3_bench_full.json').read().splitlines() if l.startswith('{')][-1])
    print('HEAD', d['config']['workload'][:40], 'value %.3e ms %.1f e2e ms %.1f roof %s frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'] or 0))
    print('kms', {k: round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items() if k!='launches'})
    print('lut', json.dumps(d.get('lut_build_s'))[:900])
    print('cpu', json.dumps(d.get('cpu_baseline'))[:700])
    print('clocks', d['clocks'])
except Exception as e:
    print('bench_full parse failed', e)
PY
echo "== reference arm"
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/p3_ref.json 2> gpurun_out/p3_ref.err ) 2>&1 | grep real
head -c 1200 gpurun_out/p3_ref.json; echo; tail -c 400 gpurun_out/p3_ref.err
echo "== ncu"
GEOSMIE_NO_SMALL=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_coeff" -s 6 -c 1 -f -o gpurun_out/r02b_coeff_default_build_su_2196cells \
   python tools/one_step.py su 2 > gpurun_out/p3_ncu1.log 2>&1; tail -2 gpurun_out/p3_ncu1.log
GEOSMIE_NO_SMALL=1 GEOSMIE_B200_LIB=tools/variants/lib_tpc4.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_coeff" -s 6 -c 1 -f -o gpurun_out/r02b_coeff_tpc4_build_su_2196cells \
   python tools/one_step.py su 2 > gpurun_out/p3_ncu2.log 2>&1; tail -2 gpurun_out/p3_ncu2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_small|k_gram$|k_gram_eval" -s 9 -c 3 -f -o gpurun_out/r02b_small_gram_eval_su_2196cells \
   python tools/one_step.py su 2 > gpurun_out/p3_ncu3.log 2>&1; tail -2 gpurun_out/p3_ncu3.log
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_su.csv \
   python tools/one_step.py su 2 > gpurun_out/p3_ncu4.log 2>&1; tail -1 gpurun_out/p3_ncu4.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_ss.csv \
   python tools/one_step.py ss 2 > gpurun_out/p3_ncu5.log 2>&1; tail -1 gpurun_out/p3_ncu5.log
python tools/dram_bytes.py gpurun_out/launches_su.csv su 2196 gpurun_out/launches_ss.csv ss 10980 > gpurun_out/r02_dram_bytes.json 2> gpurun_out/p3_dram.err; tail -3 gpurun_out/p3_dram.err; head -c 600 gpurun_out/r02_dram_bytes.json
