#!/usr/bin/env python3
"""Compare table files written by runoptics / rungsf with the golden tables of the unmodified reference (tests/golden/full_<sp>.npz).

    python tools/check_tables.py <dir> su bc ss
Prints, per species: max relative error of every scalar variable (qabs on the scale of qext), max phase-matrix error of the stratified
golden cells relative to max P11, pback, bit-equality of the bookkeeping variables, and -- when optics_<sp>.nc4 with `pmom` exists --
the largest re-synthesis error of the moments with the reference's own basis (tests/golden/gsf_basis.npz) over those cells."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from geosmie_b200 import ncio

G = os.path.join(ROOT, "tests", "golden")
EXACT = ("rLow", "rUp", "growth_factor", "rhop", "refreal", "refimag")


def check(d, sp):
    g = np.load(os.path.join(G, "full_%s.npz" % sp))
    nc = ncio.Dataset(os.path.join(d, "optics_%s.nomom.nc4" % sp), "r")
    V = {k: np.array(nc.variables[k][:]) for k in nc.variables}
    hp = sp == "bc"                       # hydrophobic species: the file has the extra first bin; the golden table is the computed bin
    res = {"species": sp, "scalars_max_rel": 0.0, "worst_scalar": None, "exact": True}
    for key in [k[5:] for k in g.files if k.startswith("var__")]:
        if key in ("rh", "wavelength", "bin", "p", "ang"):
            continue
        a, r = V[key], g["var__" + key]
        if hp:
            a = a[1:]
        if key in EXACT:
            res["exact"] &= bool(np.array_equal(a, r))
            continue
        if key == "qabs":
            e = float(np.max(np.abs(a - r) / np.maximum(np.abs(r), np.abs(V["qext"][1:] if hp else V["qext"]))))
        elif key == "pback":
            e = float(np.max(np.abs(a - r) / np.abs(r[..., :1])))
            res["pback_max_rel_to_p11back"] = e
            continue
        else:
            e = float(np.max(np.abs(a - r) / np.maximum(np.abs(r), 1e-300)))
        if e > res["scalars_max_rel"]:
            res["scalars_max_rel"], res["worst_scalar"] = e, key
    cells = g["phase_cells"]
    pe = 0.0
    for key in ("p11", "p12", "p22", "p33", "p34", "p44"):
        a = V[key][1:] if hp else V[key]
        got = np.stack([a[:, li, rhi, :] for li, rhi in cells], axis=1)
        scale = np.abs(g["phase__p11"]).max(axis=-1, keepdims=True)
        pe = max(pe, float(np.max(np.abs(got - g["phase__" + key]) / scale)))
    res["phase_max_rel_to_p11max"] = pe
    res["cells"] = int(np.prod(V["qext"].shape))
    full = os.path.join(d, "optics_%s.nc4" % sp)
    if os.path.exists(full):
        B = np.load(os.path.join(G, "gsf_basis.npz"))
        pm = np.array(ncio.Dataset(full, "r").variables["pmom"][:])        # (bin, wavelength, rh, p, m)
        worst = 0.0
        per_bin = [0.0] * pm.shape[0]
        for li, rhi in cells[::6]:
            for b in range(pm.shape[0]):
                m_ = pm[b, li, rhi]
                a2p3 = (m_[4] + m_[2]) @ B["d22"]
                a2m3 = (m_[4] - m_[2]) @ B["d2m2"]
                p22 = 0.5 * (a2p3 + a2m3)
                R = {"p11": m_[0] @ B["leg"], "p12": m_[1] @ B["d02"], "p22": p22, "p33": a2p3 - p22, "p34": m_[3] @ B["d02"], "p44": m_[5] @ B["leg"]}
                sc = np.abs(V["p11"][b, li, rhi]).max()
                for k in R:
                    e = float(np.max(np.abs(R[k] - V[k][b, li, rhi])) / sc)
                    worst = max(worst, e)
                    per_bin[b] = max(per_bin[b], e)
        res["pmom_shape"] = list(pm.shape)
        # 129 moments represent the phase matrix of particles up to x ~ 50; for the large sea-salt bins (x up to 2513, forward peak ~1e6) the
        # truncated series cannot reproduce it -- there the number measures the truncation the reference's NSPHER = 129 imposes, not an error
        res["pmom_resynthesis_max_rel_to_p11max"] = worst
        res["pmom_resynthesis_per_bin"] = per_bin
    return res


if __name__ == "__main__":
    d = sys.argv[1]
    for sp in sys.argv[2:]:
        if not os.path.exists(os.path.join(G, "full_%s.npz" % sp)):
            print(json.dumps({"species": sp, "skipped": "no golden table"}))
            continue
        print(json.dumps(check(d, sp)))
