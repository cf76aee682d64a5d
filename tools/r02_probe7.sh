#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
tools/tables_multigpu.sh $N 2>&1 | tail -5 | cut -c1-500
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29811 bench.py --gpus $N --steps 3 --warmup 3 --workloads su > gpurun_out/p7_bench_${N}gpu.json 2> gpurun_out/p7_bench_${N}gpu.err
echo "bench rc=$?"; grep -v "^$\|####\|Starting\|RADIND\|Done\|mode pygeos\|done" gpurun_out/p7_bench_${N}gpu.err | tail -8
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open('gpurun_out/p7_bench_%sgpu.json' % n).read().splitlines() if l.startswith('{')][-1])
    print('lut N=%s' % n, json.dumps({k: (round(v.get('s', -1), 3), round(v.get('runoptics_s', v.get('table_s', -1)), 3)) if isinstance(v, dict) else '' for k, v in d.get('lut_build_s').items()}))
except Exception as e:
    print('parse failed', e)
PY
export CUDA_VISIBLE_DEVICES=0
timeout 600 python bench.py --steps 3 --warmup 3 --workloads su --no-cpu-baseline > gpurun_out/p7_bench_1gpu.json 2> gpurun_out/p7_bench_1gpu.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/p7_bench_1gpu.json').read().splitlines() if l.startswith('{')][-1])
    print('lut N=1', json.dumps({k: (round(v.get('s', -1), 3), round(v.get('runoptics_s', v.get('table_s', -1)), 3)) if isinstance(v, dict) else '' for k, v in d.get('lut_build_s').items()}))
except Exception as e:
    print('parse failed', e)
PY
