#!/bin/bash
# bench.py on N GPUs only (final code): both workloads, lut_build_s, fine-grid build
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29900 + N)) bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/r02_bench_${N}gpu_final.json 2> gpurun_out/r02_bench_${N}gpu_final.err
echo "bench N=$N rc=$?"
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open('gpurun_out/r02_bench_%sgpu_final.json' % n).read().splitlines() if l.startswith('{')][-1])
    for w, r in d['workloads'].items():
        print(w, 'N=%s value %.4e ms %.3f e2e ms %.3f gather %s' % (n, r['value'], r['ms_per_step'], r['e2e']['ms_per_step'], r['gather']))
        print('   per-rank device ms', [round(x, 3) for x in r['per_rank_ms_per_step']['device']])
        print('   per-rank k_coeff', [k.get('k_coeff') for k in r['per_rank_kernel_ms']])
    print('   lut', json.dumps({k: round(v.get('s', -1), 3) for k, v in d['lut_build_s'].items() if isinstance(v, dict)}))
    print('   clocks', d['clocks'].get('per_rank_sm_mhz'), d['clocks'].get('reasons'))
except Exception as e:
    print('parse failed', e)
    print(open('gpurun_out/r02_bench_%sgpu_final.err' % n).read()[-1500:])
PY
