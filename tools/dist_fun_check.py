#!/usr/bin/env python3
"""Multi-GPU check (run under torchrun): dointegration.fun with cells sharded over the ranks and gathered with NCCL must
reproduce the reference's table for the mini species fixtures.  torchrun --nproc-per-node 2 tools/dist_fun_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import fun_fixture, run_dir
from geosmie_b200 import dist, dointegration as DI

comm = dist.Comm.from_env()
ok = True
for name in ("su_mini", "ss_mini", "mm_mini", "bc_mini"):
    for device_psd in (True, False):
        g, files, base = fun_fixture(name)
        with run_dir(files) as d:
            out = DI.fun(base + ".json", "json", d, False, write=False, comm=comm, device_psd=device_psd)
        if comm.rank == 0:
            worst = 0.0
            for k in ("qext", "qsca", "g", "bext", "bbck", "rEff", "mass", "volume", "area", "lidar_ratio"):
                worst = max(worst, float(np.max(np.abs(out["vals"][k] - g["var__" + k]) / np.abs(g["var__" + k]))))
            p = float(np.max(np.abs(out["vals"]["p11"] - g["var__p11"]) / np.abs(g["var__p11"]).max(axis=-1, keepdims=True)))
            good = worst < 1e-9 and p < 1e-7
            ok &= good
            print("%s device_psd=%d world=%d: scalars %.2e phase %.2e %s" % (name, device_psd, comm.world, worst, p, "OK" if good else "FAIL"), flush=True)
        else:
            assert out is None
comm.barrier()
comm.close()
if comm.rank == 0:
    print("DIST_FUN_OK" if ok else "DIST_FUN_FAIL")
    sys.exit(0 if ok else 1)
