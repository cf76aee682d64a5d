#!/bin/bash
# 2 GPUs: sharded table builds through the drivers + golden check; bench.py under torchrun (both workloads, lut_build_s); L2 window experiment.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
tools/tables_multigpu.sh $N 2>&1 | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29801 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/p5_bench_${N}gpu.json 2> gpurun_out/p5_bench_${N}gpu.err
echo "bench rc=$?"; tail -c 600 gpurun_out/p5_bench_${N}gpu.err
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open('gpurun_out/p5_bench_%sgpu.json' % n).read().splitlines() if l.startswith('{')][-1])
    for w, r in d['workloads'].items():
        print(w, 'value %.3e ms %.3f e2e ms %.3f gather %s per-rank %s' % (r['value'], r['ms_per_step'], r['e2e']['ms_per_step'], r['gather'], r.get('per_rank_ms_per_step')))
    print('lut', json.dumps(d.get('lut_build_s'))[:800])
except Exception as e:
    print('parse failed', e)
PY
line() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=d['roofline']['kernel_ms_per_step']
print('$1 step %.3f e2e %.3f k_coeff %.3f k_small %.3f k_gram %.3f sum+eval %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], k['k_coeff'], k.get('k_small',0), k['k_gram'], k['k_gram_sum_eval']))"; }
B="--no-cpu-baseline --no-lut --workloads su --steps 20"
export CUDA_VISIBLE_DEVICES=0
GEOSMIE_NO_SMALL=1 GEOSMIE_B200_LIB=tools/variants/lib_tpc4.so timeout 200 python bench.py $B 2>/dev/null | line tpc4_r01path
GEOSMIE_L2_PERSIST=32 GEOSMIE_NO_SMALL=1 GEOSMIE_B200_LIB=tools/variants/lib_tpc4.so timeout 200 python bench.py $B 2>/dev/null | line tpc4_r01path_L2persist32
GEOSMIE_L2_PERSIST=32 GEOSMIE_NO_SMALL=1 timeout 200 python bench.py $B 2>/dev/null | line r01path_L2persist32
GEOSMIE_L2_PERSIST=32 timeout 200 python bench.py $B 2>/dev/null | line new_L2persist32
timeout 200 python bench.py $B 2>/dev/null | line new
