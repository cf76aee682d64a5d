#!/bin/bash
# tests, smoke, the default bench line and the reference arm (no profiler passes)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err ) 2>&1 | grep real
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
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null ) 2>&1 | grep real
head -c 300 gpurun_out/r02_bench_reference_arm.json; echo
