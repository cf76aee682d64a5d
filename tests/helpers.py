"""Shared helpers for the test-suite."""
import contextlib
import json
import os
import tempfile

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def table_angles():
    return np.concatenate([np.linspace(0., 1., 100, endpoint=False), np.linspace(1., 10., 100, endpoint=False),
                           np.linspace(10., 180., 171, endpoint=True)])


def relerr(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


@contextlib.contextmanager
def run_dir(files=None, water=True):
    """Scratch run directory laid out like the reference's (CWD-relative data/ ...): writes `files` and, from the
    golden fixture, data/refrac.water.txt (13 header lines, columns wavelength[um] n k: particleparams.py:79-82)."""
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "data"))
        if water:
            w = load("hostlogic.npz")["su__water"]
            with open(os.path.join(d, "data", "refrac.water.txt"), "w") as fp:
                fp.write("# header\n" * 13)
                for i in range(w.shape[1]):
                    um = float("%.10g" % (w[0, i] * 1e6))       # the table's own decimal value, so that um * 1e-6 is bit-exact
                    assert um * 1e-6 == w[0, i]
                    fp.write("%.17g %.17g %.17g\n" % (um, w[1, i], w[2, i]))
        for name, text in (files or {}).items():
            with open(os.path.join(d, name), "w") as fp:
                fp.write(text)
        os.chdir(d)
        try:
            yield d
        finally:
            os.chdir(old)


def fun_fixture(name):
    """A tests/golden/fun_<name>.npz fixture: (golden dict, files to write into the run dir)."""
    g = load("fun_%s.npz" % name)
    base = name.replace("_legacy", "")
    files = dict(json.loads(str(g["files_json"])))
    files[base + ".json"] = str(g["config_json"])
    return g, files, base


def species_files(sp):
    """Files for a run dir that reproduce a shipped species config (su / ss / bc) from the recorded fixture: <sp>.json with
    the parameters of src/config/geosparticles/<sp>.json and its refractive-index table re-written in 'wsv' format."""
    from geosmie_b200.workloads import SPECIES
    ml = load("hostlogic.npz")[sp + "__mlist"]
    lines = []
    for i in range(ml.shape[1]):
        um = float("%.10g" % (ml[0, i] * 1e6))
        assert um * 1e-6 == ml[0, i]
        lines.append("%.17g %.17g %.17g" % (um, ml[1, i], ml[2, i]))
    cfg = json.loads(json.dumps(SPECIES[sp]))
    cfg["ri"] = {"format": "wsv", "path": ["ri-%s.wsv" % sp]}
    return {sp + ".json": json.dumps(cfg), "ri-%s.wsv" % sp: "\n".join(lines) + "\n"}
