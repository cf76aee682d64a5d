#!/bin/bash
# A/B of a host-side file path switch (here: copy_file_range for the untouched variables) in one call, runs alternating
set -u
cd "$(dirname "$0")/.."
cat > /tmp/ab.py <<'PY'
import contextlib, io, os, sys, tempfile, time
sys.path.insert(0, os.getcwd())
from geosmie_b200 import runoptics, workloads
from geosmie_b200.gsf import rungsf
sp = sys.argv[1]
with tempfile.TemporaryDirectory() as d:
    cfg = workloads.write_run_dir(d, sp); os.chdir(d)
    for attempt in range(11):
        old = attempt % 2 == 1
        if old: os.environ["GEOSMIE_NO_COPY_FILE_RANGE"] = "1"
        else: os.environ.pop("GEOSMIE_NO_COPY_FILE_RANGE", None)
        out = os.path.join(d, "o%d" % attempt); os.makedirs(out)
        with contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter(); runoptics.main(["--name", cfg, "--dest", out]); t1 = time.perf_counter()
            rungsf.main(["--filename", os.path.join(out, "optics_%s.nomom.nc4" % sp), "--dest", out]); t2 = time.perf_counter()
        print("%s %s run %d: runoptics %.3f s, rungsf %.3f s, total %.3f" % (sp, "OLD" if old else "NEW", attempt, t1 - t0, t2 - t1, t2 - t0))
PY
python /tmp/ab.py ss
python /tmp/ab.py su
