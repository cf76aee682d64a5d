#!/usr/bin/env python3
"""Diagnostic: table builds after a dense optics_SS benchmark in the same process (pool / allocator state)."""
import cProfile, contextlib, io, os, pstats, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench
from geosmie_b200 import _lib, runoptics, workloads
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
h = _lib.Handle.get(0)
stream = torch.cuda.current_stream()
h.set_stream(stream.cuda_stream)
tb = bench.TableBench("ss", torch, dev, h, stream, 1, 0)
for _ in range(2):
    tb.step_device()
tb.prepare_e2e()
tb.step_e2e()
tb.finish_e2e()
torch.cuda.synchronize()
tb.close()
print("bench part done", flush=True)
with tempfile.TemporaryDirectory() as d:
    cfg = workloads.write_run_dir(d, "ss")
    os.chdir(d)
    for attempt in range(4):
        out = os.path.join(d, "o%d" % attempt); os.makedirs(out)
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            pr.enable()
            runoptics.main(["--name", cfg, "--dest", out])
            pr.disable()
        print("run %d: runoptics %.3f s" % (attempt, time.perf_counter() - t0), flush=True)
        if attempt in (1, 3):
            s = io.StringIO()
            pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14)
            print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:]))
