#!/usr/bin/env python3
"""60-digit (mpmath) truth for the coated-BC TABLE cells of tests/test_gpu_parity.py::test_coated_table_against_oracle.

The oracle (like the reference) evaluates coated_mie_coeff with scipy's complex-argument jv/yv, which limits ITS accuracy to ~1e-8
on some particles; with this fixture the CUDA table path is held to the north-star tolerance (1e-9 on the scalar sums, 1e-7 on the
phase sums) against the exactly evaluated formula instead of a loosened tolerance against the noisy one.
    python tests/golden/make_truth_table.py [nproc]     # ~15 min of two cores -> tests/golden/coated_table_truth.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
CELLS = [(3, 0), (3, 20), (30, 35), (55, 28)]
ANG_STRIDE = 6


def _angles():
    return np.concatenate([np.linspace(0., 1., 100, endpoint=False), np.linspace(1., 10., 100, endpoint=False),
                           np.linspace(10., 180., 171, endpoint=True)])


def one(args):
    import mpmath as mp
    from make_truth import coated_mp
    from oracle import mie_oracle as mo
    mp.mp.dps = 60
    x, y, m1, m2 = args
    if y == x:                       # no shell: homogeneous sphere of the core material, evaluated through the same formula with m2 = m1
        an, bn, nmax = coated_mp(mp.mpc(m1 ** 2), mp.mpc(m1 ** 2), mp.mpf(x) * mp.mpf("0.5"), mp.mpf(y))
    else:
        an, bn, nmax = coated_mp(mp.mpc(m1 ** 2), mp.mpc(m2 ** 2), mp.mpf(x), mp.mpf(y))
    q = mo.mie_props(an, bn, nmax, y)
    u = np.cos(np.radians(_angles()[::ANG_STRIDE]))
    mu = np.zeros((6, u.size))
    for a, uu in enumerate(u):
        s1, s2 = mo.mie_s12(an, bn, nmax, uu)
        mu[0, a] = 0.5 * (abs(s1) ** 2 + abs(s2) ** 2)
        mu[1, a] = 0.5 * (abs(s2) ** 2 - abs(s1) ** 2)
        mu[3, a] = (s1 * np.conj(s2)).real
        mu[4, a] = -(np.conj(s1) * s2).imag
    return q, mu


if __name__ == "__main__":
    import multiprocessing as mp_
    from geosmie_b200 import coated_table, workloads
    from oracle import mie_oracle as mo
    params, lambarr, part_m, water_m, _ = workloads.species_inputs("bc")
    xx, cl = coated_table.coated_cells(params, lambarr, part_m, water_m, cells=CELLS)
    out = {"cells": np.array(CELLS)}
    with mp_.get_context("fork").Pool(int(sys.argv[1]) if len(sys.argv) > 1 else 2) as pool:
        for ci, c in enumerate(cl):
            nz = np.where(c['w'] > 0)[0]
            y = c['gf'] * xx
            res = pool.map(one, [(float(xx[i]), float(y[i]), complex(c['m1']), complex(c['m2'])) for i in nz], chunksize=4)
            q = np.zeros((xx.size, 6))
            mu = np.zeros((6, xx.size, len(_angles()[::ANG_STRIDE])))
            for i, (qi, mi) in zip(nz, res):
                q[i] = qi
                mu[:, i, :] = mi
            s_o, p_o = mo.raw_sums(y, q, mu, c['w'])
            out["scal_%d" % ci], out["phase_%d" % ci] = s_o, p_o
            print("cell", c['cell'], "particles", nz.size, s_o[:5], flush=True)
    np.savez_compressed(os.path.join(HERE, "coated_table_truth.npz"), **out)
    print("wrote coated_table_truth.npz")
