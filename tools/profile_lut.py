#!/usr/bin/env python3
"""Where the wall time of a table build goes: cProfile of runoptics.main + rungsf.main for one species (second run, warm).
    python tools/profile_lut.py ss"""
import cProfile, contextlib, io, os, pstats, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from geosmie_b200 import runoptics, workloads
from geosmie_b200.gsf import rungsf
sp = sys.argv[1] if len(sys.argv) > 1 else "ss"
with tempfile.TemporaryDirectory() as d:
    cfg = workloads.write_run_dir(d, sp)
    os.chdir(d)
    for attempt in range(2):
        out = os.path.join(d, "o%d" % attempt); os.makedirs(out)
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            pr.enable()
            runoptics.main(["--name", cfg, "--dest", out])
            t1 = time.perf_counter()
            rungsf.main(["--filename", os.path.join(out, "optics_%s.nomom.nc4" % sp), "--dest", out])
            pr.disable()
        t2 = time.perf_counter()
        print("%s run %d: runoptics %.3f s, rungsf %.3f s" % (sp, attempt, t1 - t0, t2 - t1))
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:]))
