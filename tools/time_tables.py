#!/usr/bin/env python3
"""LUT build timing: optics_SU / optics_SS / optics_BC (+ GSF moments) on one GPU, host and device parts separately.

    python tools/time_tables.py [su ss bc] [--dense]
Prints one JSON line per species.  (BASELINE.json metric: 'optics_XX LUT build time'.)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from geosmie_b200 import _lib, dointegration as DI, workloads  # noqa: E402
from geosmie_b200.gsf import convertncdf  # noqa: E402


def build(sp, dense):
    h = _lib.Handle.get(0)
    ang = DI.table_angles()
    cost = np.cos(np.radians(ang))
    t_host = t_gpu = t_gsf = 0.0
    evals = 0
    kms = {"coeff": 0.0, "contract": 0.0, "finalize": 0.0}
    stats_tot = {}
    rets = []
    for b in range(workloads.n_bins(sp)):
        t0 = time.time()
        plan = workloads.bin_plan(sp, b, device_psd=DEVICE_PSD)
        t1 = time.time()
        table = _lib.Table(plan.xx, plan.nmax, cost, h)
        table.set_timing(True)
        scal, phase, tpc = plan.evaluate(table, elide=not dense)
        t2 = time.time()
        k = table.last_kernel_ms()
        st = table.last_stats()
        for kk in kms:
            kms[kk] += k[kk]
        for kk, v in st.items():
            stats_tot[kk] = stats_tot.get(kk, 0) + v
        ret = DI.postprocess(plan.reduce(scal, phase, tpc), ang)
        t3 = time.time()
        F = np.stack([ret[k2] for k2 in convertncdf.MISH_KEYS], axis=1)
        coef, _ = h.gsf_expand(ang, F, 129, quantize10=True)
        t4 = time.time()
        table.close()
        t_host += (t1 - t0) + (t3 - t2)
        t_gpu += t2 - t1
        t_gsf += t4 - t3
        evals += len(plan.cells) * plan.xx.size
        rets.append((ret, coef))
    return {"species": sp, "dense": dense, "device_psd": DEVICE_PSD, "grid_particle_evals": evals, "host_inputs_s": t_host, "gpu_call_s": t_gpu, "gsf_s": t_gsf,
            "total_s": t_host + t_gpu + t_gsf, "kernel_ms": kms, "stats": stats_tot,
            "grid_evals_per_s_gpu_call": evals / t_gpu}


DEVICE_PSD = "--host-psd" not in sys.argv

def build_coated():
    """BASELINE config 4 (extension): coated BC table, 615 sizes x 61 lambda x 36 RH = 1.35 M coated particle evaluations."""
    from geosmie_b200 import coated_table
    params, lambarr, part_m, water_m, _ = workloads.species_inputs("bc")
    coated_table.build(params, lambarr, part_m, water_m, cells=[(0, 0), (0, 35)])     # warm-up (module load, allocations)
    t0 = time.time()
    cells, ret, _ = coated_table.build(params, lambarr, part_m, water_m, elide="--dense" not in sys.argv)
    dt = time.time() - t0
    return {"species": "bc_coated", "dense": "--dense" in sys.argv, "grid_particle_evals": len(cells) * 615, "total_s": dt,
            "grid_evals_per_s": len(cells) * 615 / dt, "cells": len(cells)}


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")] or ["su", "bc", "ss"]
    dense = "--dense" in sys.argv
    for sp in args:
        print(json.dumps(build_coated() if sp == "bc_coated" else build(sp, dense)), flush=True)
