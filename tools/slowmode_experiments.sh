#!/bin/bash
# One multi-GPU gpurun call that runs every staged experiment against the rank-dependent slow mode of k_coeff
# and prints one line per variant: step time, per-rank step times, per-rank k_coeff times, lowest clock / highest power seen.
#
#   (here)      make -C geosmie_b200/csrc && tools/slowmode_experiments.sh build      # variant library with -DGM_COEFF_TPC=4
#   (GPU box)   gpurun --gpus 4 --timeout 400 -- 'tools/slowmode_experiments.sh run 4'
#   (GPU box)   gpurun --timeout 400 -- 'tools/slowmode_experiments.sh perf'          # 1 GPU: parity + step time of the k_coeff variants
set -u
cd "$(dirname "$0")/.."
if [ "${1:-}" = "build" ]; then
  mkdir -p tools/variants
  (cd geosmie_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
     -DGM_COEFF_TPC=4 -o ../../tools/variants/lib_tpc4.so gm_api.cu gm_gsf.cu gm_bands.cu gm_peer.cu) && echo "built tools/variants/lib_tpc4.so"
  (cd geosmie_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
     -DGM_COEFF_SHORT_START=1 -o ../../tools/variants/lib_short.so gm_api.cu gm_gsf.cu gm_bands.cu gm_peer.cu) && echo "built tools/variants/lib_short.so"
  exit $?
fi
if [ "${1:-}" = "perf" ]; then
  # 1 GPU: parity of the k_coeff variants (whole GPU suite through the variant library) and their step times
  for L in geosmie_b200/libgeosmie_b200.so tools/variants/lib_tpc4.so tools/variants/lib_short.so; do
    [ -f $L ] || continue
    echo "== $L"
    GEOSMIE_B200_LIB=$L timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
    GEOSMIE_B200_LIB=$L timeout 120 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
print('step %.3f e2e %.3f k_coeff %.3f k_gram %.3f sum+eval %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], k['k_coeff'], k['k_gram'], k['k_gram_sum_eval']))"
  done
  exit 0
fi
N=${2:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29600
run() {   # name, env assignments...
  local name=$1; shift
  port=$((port + 1))
  env "$@" timeout 120 $TR --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline \
      > gpurun_out/slow_$name.json 2> gpurun_out/slow_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    txt = open("gpurun_out/slow_%s.json" % name).read()
    d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    c = d.get("clocks") or {}
    print("%-12s step %.3f ms  per-rank %s  k_coeff %s  e2e %.3f  sm_mhz_min %s  power_w_max %s  brake %s" % (
        name, d["ms_per_step"], [round(x, 2) for x in d["per_rank_ms_per_step"]["device"]],
        [k["k_coeff"] for k in d["per_rank_kernel_ms"]], d["e2e"]["ms_per_step"], c.get("per_rank_sm_mhz_min"),
        c.get("per_rank_power_w_max"), c.get("hw_power_brake")), flush=True)
except Exception as e:   # noqa: BLE001
    print("%-12s FAILED: %s" % (name, e), flush=True)
PY
}
run baseline   GEOSMIE_GATHER=peer
run gloo       GEOSMIE_GATHER=peer GEOSMIE_BENCH_CONTROL=gloo
run stagger    GEOSMIE_GATHER=peer GEOSMIE_BENCH_STAGGER_MS=3
run nvls_off   GEOSMIE_GATHER=peer NCCL_NVLS_ENABLE=0
[ -f tools/variants/lib_tpc4.so ] && run tpc4 GEOSMIE_GATHER=peer GEOSMIE_B200_LIB=tools/variants/lib_tpc4.so
run baseline2  GEOSMIE_GATHER=peer
