#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29951 tools/profile_fine.py ss 2048 2>&1 | grep -v "^$\|OMP_NUM\|\*\*\*\*" | head -45 | cut -c1-170
CUDA_VISIBLE_DEVICES=0 python tools/profile_fine.py ss 2048 2>&1 | head -24 | cut -c1-170
