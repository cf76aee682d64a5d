#!/usr/bin/env python3
"""cProfile of the BASELINE config-5 table build (dointegration.fun on 2048 wavelengths, no phase matrices kept): where the host time goes.
    python tools/profile_fine.py [ss] [nlam]"""
import cProfile, contextlib, io, os, pstats, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from geosmie_b200 import dointegration as DI, workloads
sp = sys.argv[1] if len(sys.argv) > 1 else "ss"
nlam = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
files, cfg = workloads.fine_grid_files(sp, nlam)
world = int(os.environ.get("WORLD_SIZE", "1"))
comm = None
if world > 1:
    from geosmie_b200 import dist
    comm = dist.Comm.from_env()
rank = 0 if comm is None else comm.rank
with tempfile.TemporaryDirectory() as d:
    for name, text in files.items():
        os.makedirs(os.path.dirname(os.path.join(d, name)), exist_ok=True)
        open(os.path.join(d, name), "w").write(text)
    os.chdir(d)
    for attempt in range(4):
        pr = cProfile.Profile()
        if comm is not None:
            comm.barrier()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            pr.enable()
            out = DI.fun(cfg, "json", d, False, write=False, keep_phase=False, comm=comm)
            pr.disable()
        if rank == 0:
            print("%s %d wavelengths on %d GPU(s), run %d: %.3f s" % (sp, nlam, world, attempt, time.perf_counter() - t0))
    if rank == 0:
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(30)
        print("\n".join(l[:160] for l in s.getvalue().splitlines()[4:]))
if comm is not None:
    comm.close()
