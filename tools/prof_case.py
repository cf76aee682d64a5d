#!/usr/bin/env python3
"""Run one gm_table_run on a slice of a species bin (for ncu captures): python tools/prof_case.py ss 4 24 [--elide]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from geosmie_b200 import _lib, dointegration as DI, workloads
sp, b, ncell = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
elide = "--elide" in sys.argv
cells = [(li, rhi) for li in range(0, 61, 7) for rhi in range(0, 36, 5)][:ncell]
if ncell > len(cells):
    cells = [(li, rhi) for li in range(61) for rhi in range(36)][:ncell]
plan = workloads.bin_plan(sp, b, cells=cells)
h = _lib.Handle.get(0)
cost = np.cos(np.radians(DI.table_angles()))
t = _lib.Table(plan.xx, plan.nmax, cost, h)
t.set_timing(True)
mz, wp, ws, tpc = plan.tasks()
for it in range(3):
    t0 = time.time(); scal, phase = t.run(mz, mz, wp, ws, elide=elide); dt = time.time() - t0
print(sp, b, len(plan.cells), "cells", "call %.1f ms" % (dt * 1e3), t.last_kernel_ms(), t.last_stats())
