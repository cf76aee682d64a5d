#!/bin/bash
# Refresh of the optics_SU evidence after the Chebyshev-node evaluation: default bench line, ncu launch list of bench.py (SU), DRAM bytes (SU), ncu --set full (SU).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_1gpu.json').read().splitlines() if l.startswith('{')][-1])
for w, r in d['workloads'].items():
    k = r['roofline']['kernel_ms_per_step']
    print(w, 'value %.4e ms %.3f e2e %.3f roof %s frac %.3f traffic %s' % (r['value'], r['ms_per_step'], r['e2e']['ms_per_step'], r['roofline']['kernel'], r['roofline']['frac'], r['roofline']['traffic']))
    print('   ', {a: round(b, 3) for a, b in k.items() if a.startswith('k_')})
print('lut', json.dumps({k: (round(v.get('s', -1), 3)) for k, v in d['lut_build_s'].items() if isinstance(v, dict)}))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['kind'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_su.csv \
   python bench.py --workloads su --steps 2 --warmup 3 --no-lut --no-cpu-baseline > gpurun_out/f_ncu_su.log 2>&1; wc -l gpurun_out/r02_launches_bench_su.csv
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_su.csv \
   python tools/one_step.py su 2 > gpurun_out/f_ncu4.log 2>&1; tail -1 gpurun_out/f_ncu4.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_small|k_coeff|k_gram$|k_gram_eval|k_gram_sum|k_gram_interp" -s 28 -c 7 -f -o gpurun_out/r02h_su_step_kernels \
   python tools/one_step.py su 2 > gpurun_out/f_ncu6.log 2>&1; grep -c "Profiling" gpurun_out/f_ncu6.log
