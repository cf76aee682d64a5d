#!/usr/bin/env python3
"""One warm-up step + one measured dense step of a workload (device-pointer path incl. the GSF stage), for ncu launch lists:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X \
        python tools/one_step.py su|ss [nsteps]
tools/dram_bytes.py turns the CSV into profiles/r02_dram_bytes.json (per kernel and per cell)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench
from geosmie_b200 import _lib
sp = sys.argv[1] if len(sys.argv) > 1 else "su"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
h = _lib.Handle.get(0)
stream = torch.cuda.current_stream()
h.set_stream(stream.cuda_stream)
tb = bench.TableBench(sp, torch, dev, h, stream, 1, 0)
torch.cuda.synchronize()
print("MARK setup done", flush=True)
for _ in range(nsteps):
    tb.step_device()
    torch.cuda.synchronize()
print("cells", tb.cells, "evals", tb.evals, "kernel ms", tb.kernel_ms())
tb.close()
