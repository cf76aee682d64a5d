#!/bin/bash
# Final measurements of the round on one GPU: tests, smoke, the default bench line, the reference arm, ncu launch lists of bench.py itself,
# DRAM bytes per kernel, ncu --set full of the main kernels of the final revision.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_1gpu.json').read().splitlines() if l.startswith('{')][-1])
for w, r in d['workloads'].items():
    k = r['roofline']['kernel_ms_per_step']
    print(w, 'value %.4e ms %.3f e2e %.3f roof %s frac %.3f traffic %s' % (r['value'], r['ms_per_step'], r['e2e']['ms_per_step'], r['roofline']['kernel'], r['roofline']['frac'], r['roofline']['traffic']))
    print('   ', {a: round(b, 3) for a, b in k.items() if a.startswith('k_')})
print('lut', json.dumps({k: (round(v.get('s', -1), 3)) for k, v in d['lut_build_s'].items() if isinstance(v, dict)}))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['kind'], '| port', d['cpu_baseline']['port']['value'])
print('clocks', d['clocks'])
PY
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null ) 2>&1 | grep real
head -c 400 gpurun_out/r02_bench_reference_arm.json; echo
timeout 300 python tools/profile_fine.py ss 2048 > gpurun_out/f_profile_fine.txt 2>&1; head -40 gpurun_out/f_profile_fine.txt
echo "== ncu launch lists of bench.py"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_su.csv \
   python bench.py --workloads su --steps 2 --warmup 3 --no-lut --no-cpu-baseline > gpurun_out/f_ncu_su.log 2>&1; wc -l gpurun_out/r02_launches_bench_su.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_bench_ss.csv \
   python bench.py --workloads ss --steps 2 --warmup 3 --no-lut --no-cpu-baseline > gpurun_out/f_ncu_ss.log 2>&1; wc -l gpurun_out/r02_launches_bench_ss.csv
echo "== DRAM bytes"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_su.csv \
   python tools/one_step.py su 2 > gpurun_out/f_ncu4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_ss.csv \
   python tools/one_step.py ss 2 > gpurun_out/f_ncu5.log 2>&1
python tools/dram_bytes.py gpurun_out/launches_su.csv su 2196 gpurun_out/launches_ss.csv ss 10980 > gpurun_out/r02_dram_bytes.json 2> gpurun_out/f_dram.err; tail -2 gpurun_out/f_dram.err
echo "== ncu --set full"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_small|k_coeff|k_gram$|k_gram_eval" -s 20 -c 5 -f -o gpurun_out/r02f_su_step_kernels \
   python tools/one_step.py su 2 > gpurun_out/f_ncu6.log 2>&1; grep -c "Profiling" gpurun_out/f_ncu6.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_coeff|k_contract" -s 2 -c 2 -f -o gpurun_out/r02f_coeff_contract_ss_bin5 \
   python tools/prof_case.py ss 4 65 > gpurun_out/f_ncu7.log 2>&1; grep -c "Profiling" gpurun_out/f_ncu7.log
