#!/bin/bash
# optics_SU / optics_BC / optics_SS regenerated on N GPUs through the drivers themselves (torchrun -m geosmie_b200.runoptics, then
# rungsf), wall times, and every table compared with the golden tables of the unmodified reference (tools/check_tables.py).
#   gpurun --gpus 8 --timeout 600 -- 'tools/tables_multigpu.sh 8'   -> gpurun_out/tables_<N>gpu.txt
set -u
cd "$(dirname "$0")/.."
ROOT=$(pwd)
N=${1:-8}
OUT=$ROOT/gpurun_out/tables_${N}gpu.txt
mkdir -p gpurun_out
D=$(mktemp -d)
python - "$D" <<'PY'
import sys
sys.path.insert(0, ".")
from geosmie_b200 import workloads
for sp in ("su", "bc", "ss"):
    workloads.write_run_dir(sys.argv[1], sp)
PY
cd "$D"
export PYTHONPATH=$ROOT
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1) x $N, $(date -u +%FT%TZ)"
port=29700
for rep in $(seq 1 ${REPS:-2}); do
for sp in su bc ss; do
  mkdir -p out_$rep
  port=$((port + 1))
  t0=$(date +%s%N)
  if [ "$N" -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port -m geosmie_b200.runoptics --name $sp.json --dest out_$rep > log_$sp.txt 2>&1
  else
    python -m geosmie_b200.runoptics --name $sp.json --dest out_$rep > log_$sp.txt 2>&1
  fi
  rc=$?
  t1=$(date +%s%N)
  python -m geosmie_b200.gsf.rungsf --filename out_$rep/optics_$sp.nomom.nc4 --dest out_$rep >> log_$sp.txt 2>&1
  t2=$(date +%s%N)
  echo "run $rep optics_$sp on $N GPU(s): runoptics rc=$rc $(( (t1 - t0) / 1000000 )) ms (process start, CUDA / NCCL initialisation, table build, file), rungsf $(( (t2 - t1) / 1000000 )) ms"
  [ $rc -ne 0 ] && tail -5 log_$sp.txt
done
done
python $ROOT/tools/check_tables.py out_${REPS:-2} su bc ss
} 2>&1 | tee $OUT
cd $ROOT; rm -rf "$D"
