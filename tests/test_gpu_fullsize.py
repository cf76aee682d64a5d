"""GPU parity at the BASELINE sizes (-m gpu): complete optics_BC / optics_SU tables and 30 cells of optics_SS against
golden vectors produced by running the unmodified reference on the shipped configs (tests/golden/make_golden_full.py)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, load, relerr, run_dir, species_files

pytestmark = pytest.mark.gpu
TOL_Q, TOL_P = 1e-9, 1e-7


def _check_full_table(sp):
    from geosmie_b200 import dointegration as DI
    g = load("full_%s.npz" % sp)
    with run_dir(species_files(sp)) as d:
        out = DI.fun(sp + ".json", "json", d, False, write=False)
    vals = out["vals"]
    for key in [k[5:] for k in g.files if k.startswith("var__")]:
        if key in ("rh", "wavelength", "bin", "p", "ang"):
            continue
        a, r = vals[key], g["var__" + key]
        assert a.shape == r.shape, key
        if key in ("rLow", "rUp", "growth_factor", "rhop", "refreal", "refimag"):
            assert np.array_equal(a, r), key
        elif key == "qabs":
            assert np.max(np.abs(a - r) / np.maximum(np.abs(r), np.abs(vals["qext"]))) < TOL_Q, key
        elif key == "pback":
            assert np.max(np.abs(a - r) / np.abs(r[..., :1])) < TOL_P, key
        else:
            assert relerr(a, r) < TOL_Q, (key, relerr(a, r))
    cells = g["phase_cells"]
    for key in ("p11", "p12", "p22", "p33", "p34", "p44"):
        ref = g["phase__" + key]                                                   # [bin, cell, ang]
        got = np.stack([vals[key][:, li, rhi, :] for li, rhi in cells], axis=1)
        scale = np.abs(g["phase__p11"]).max(axis=-1, keepdims=True)
        assert np.max(np.abs(got - ref) / scale) < TOL_P, key
    assert np.array_equal(out["wavelength"], g["var__wavelength"]) and np.array_equal(out["rh"], g["var__rh"])


def test_full_optics_bc_table():
    """optics_BC: 615 sizes x 61 wavelengths x 36 RH (x down to 1.6e-5 -- bulk values stay within 1e-9)."""
    _check_full_table("bc")


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN, "full_su.npz")), reason="full_su.npz not generated")
def test_full_optics_su_table():
    """optics_SU (BASELINE config 2): 4459 sizes x 61 wavelengths x 36 RH = 9.79 M particle evaluations."""
    _check_full_table("su")


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN, "full_ss.npz")), reason="full_ss.npz not generated")
def test_full_optics_ss_table():
    """optics_SS (BASELINE config 3), ALL 5 x 61 x 36 = 10,980 cells: every scalar variable, pback and the phase matrices of 54
    stratified cells per bin against the unmodified reference evaluated densely on the full grids (x up to 2513, nmax up to 2567;
    tests/golden/make_golden_ss_full.py, ~11 h of one core)."""
    _check_full_table("ss")


def test_optics_ss_cells():
    """optics_SS (BASELINE config 3): 6 cells of each of the 5 bins (x up to 2513, nmax up to 2567) against the
    reference's rawMie + integratePSD."""
    from geosmie_b200 import dointegration as DI, workloads
    g = load("full_ss_cells.npz")
    cells = [tuple(c) for c in g["cells"]]
    cost = np.cos(np.radians(DI.table_angles()))
    for b in range(5):
        plan = workloads.bin_plan("ss", b, cells=cells)
        ret, table = DI.run_bin(plan, cost, elide=True)
        if b == 4:
            # reference-equivalent (dense) evaluation of the largest bin: zero-weight particles contribute exact zeros
            ret_dense, _ = DI.run_bin(plan, cost, elide=False, table=table)
            for k in ("qext", "qsca", "g", "p11", "p34"):
                assert np.max(np.abs(ret_dense[k] - ret[k])) / np.abs(ret[k]).max() < 1e-12, k
        table.close()
        for ci, (li, rhi) in enumerate(plan.cells):
            key = "b%d_l%d_r%d" % (b, li, rhi)
            for k in ("qext", "qsca", "qb", "g", "csca", "cext", "bsca", "bext", "bbck", "area", "volume", "mass", "rEff", "rMass"):
                assert relerr(ret[k][ci], g[key + "__" + k]) < TOL_Q, (key, k, relerr(ret[k][ci], g[key + "__" + k]))
            assert abs(ret["qabs"][ci] - g[key + "__qabs"]) / g[key + "__qext"] < TOL_Q
            p11 = g[key + "__p11"]
            for k in ("p11", "p12", "p33", "p34"):
                assert np.max(np.abs(ret[k][ci] - g[key + "__" + k])) / np.abs(p11).max() < TOL_P, (key, k)
