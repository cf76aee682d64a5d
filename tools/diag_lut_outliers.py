#!/usr/bin/env python3
"""Diagnostic: where an outlier run of runoptics.main (optics_SS, warm) spends its time.  Profiles every run, prints the top of the
profile of the slowest and of the median run.  Outputs are deleted after each run unless KEEP=1."""
import cProfile, contextlib, io, os, pstats, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from geosmie_b200 import runoptics, workloads
from geosmie_b200.gsf import rungsf
sp = sys.argv[1] if len(sys.argv) > 1 else "ss"
base = sys.argv[2] if len(sys.argv) > 2 else None
runs = []
with tempfile.TemporaryDirectory(dir=base) as d:
    cfg = workloads.write_run_dir(d, sp)
    os.chdir(d)
    for attempt in range(12):
        out = os.path.join(d, "o%d" % attempt); os.makedirs(out)
        pr = cProfile.Profile()
        with contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter()
            pr.enable()
            runoptics.main(["--name", cfg, "--dest", out])
            pr.disable()
            t1 = time.perf_counter()
            rungsf.main(["--filename", os.path.join(out, "optics_%s.nomom.nc4" % sp), "--dest", out])
            t2 = time.perf_counter()
        runs.append((t1 - t0, attempt, pr))
        print("run %d: runoptics %.3f s rungsf %.3f s" % (attempt, t1 - t0, t2 - t1), flush=True)
        if not os.environ.get("KEEP"):
            shutil.rmtree(out, ignore_errors=True)
runs = sorted(runs[1:], key=lambda r: r[0])
for label, (t, k, pr) in (("slowest", runs[-1]), ("median", runs[len(runs) // 2])):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(12)
    print("---- %s warm run (%d, %.3f s), by own time" % (label, k, t))
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:]))
