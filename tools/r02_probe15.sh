#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 600 python bench.py --steps 3 --warmup 3 --workloads su --no-cpu-baseline > gpurun_out/p15_bench.json 2> gpurun_out/p15_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/p15_bench.json').read().splitlines() if l.startswith('{')][-1])
print(json.dumps({k: {a: round(b, 3) for a, b in v.items() if isinstance(b, float)} for k, v in d['lut_build_s'].items() if isinstance(v, dict)}))
PY
for sp in su ss; do timeout 300 python tools/profile_lut.py $sp 2>&1 | head -16 | cut -c1-150; done
