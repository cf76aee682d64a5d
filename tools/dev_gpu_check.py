"""Development probe (not a test): prints the worst deviations of the CUDA path from the oracle / goldens."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from geosmie_b200 import _lib
from geosmie_b200.pymiecoated import Mie
from oracle import mie_oracle as mo, gsf_oracle as go

G = os.path.join(ROOT, "tests", "golden")
h = _lib.Handle.get(0)

def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))

# ---- single-particle goldens
d = np.load(os.path.join(G, "mie_single.npz"))
par, us, qg, sg = d["par"], d["us"], d["q"], d["s12"]
worst = {}
for i in range(par.shape[0]):
    x, y = par[i, 0], par[i, 1]
    kw = dict(x=x, eps=complex(par[i, 2], par[i, 3]), mu=complex(par[i, 4], par[i, 5]))
    kind = "homog"
    if not np.isnan(y):
        kw.update(y=y, eps2=complex(par[i, 6], par[i, 7])); kw.pop("mu"); kind = "coated"
    elif kw["mu"] != 1: kind = "magnetic"
    m = Mie(**kw)
    q = np.array([m.qext(), m.qsca(), m.qabs(), m.qb(), m.asy(), m.qratio()])
    s1, s2 = m.S12_array(us)
    sref = sg[i]
    smax = np.abs(sref).max()
    es = max(np.abs(s1 - (sref[:, 0] + 1j * sref[:, 1])).max(), np.abs(s2 - (sref[:, 2] + 1j * sref[:, 3])).max()) / smax
    eq = rel(q, qg[i])
    w = worst.setdefault(kind, [0, 0, None])
    if eq > w[0]: w[0] = eq; w[2] = (x, y)
    w[1] = max(w[1], es)
print("single goldens: worst rel err (q, s12/max|s|, where):", worst)

# ---- size range (DMMA per-particle path) vs goldens
d = np.load(os.path.join(G, "size_range.npz"))
from geosmie_b200.pymiecoated.mie_coated import MultipleMie
for ci in range(int(d["ncase"])):
    x = d["x_%d" % ci]; mr, mi = d["m_%d" % ci]
    mm = MultipleMie(x, None, d["cost"]); mm.preCalculate()
    q, s12 = mm.calculateS12SizeRangeArrays(mr, mi)
    qg, sg = d["q_%d" % ci], d["s12_%d" % ci]
    big = x >= 0.01
    eq = rel(q[big][:, :5], qg[big][:, :5])
    es = np.max(np.abs(s12 - sg).max(axis=(1, 2)) / np.abs(sg).max(axis=(1, 2)))
    print("size_range case %d (nx=%d, m=%g+%gi): q rel %.2e (x>=0.01), s12 %.2e; small-x q rel %.2e" % (ci, x.size, mr, mi, eq, es, rel(q[~big][:, :2], qg[~big][:, :2]) if (~big).any() else 0))

# ---- table cells vs oracle raw sums
ang = np.concatenate([np.linspace(0, 1, 100, endpoint=False), np.linspace(1, 10, 100, endpoint=False), np.linspace(10, 180, 171)])
cost = np.cos(np.radians(ang))
for (lo, hi, nx) in ((5e-3, 60.0, 700), (0.5, 900.0, 333)):
    x = np.geomspace(lo, hi, nx)
    sr = mo.SizeRange(x, cost)
    t = _lib.Table(x, sr.nmax, cost, h)
    rng = np.random.default_rng(3)
    ms = [(1.43, 1e-8), (1.75, 0.44), (1.33, 0.0)]
    W = np.zeros((len(ms), nx))
    for k in range(len(ms)):
        w = np.exp(-0.5 * ((np.log(x) - np.log(3.0 * (k + 1))) / 0.8) ** 2); w[x < 4 * lo] = 0; w[x > hi / 3] = 0
        W[k] = w / w.sum()
    mz = np.array([np.sqrt(complex(a, b) ** 2 * 1.0) for a, b in ms]); mrel = np.array([np.sqrt(complex(a, b) ** 2 / 1.0) for a, b in ms])
    for elide in (False, True):
        t0 = time.time()
        scal, phase = t.run(mz, mrel, W, elide=elide)
        dt = time.time() - t0
        for k, (mr, mi) in enumerate(ms):
            q, _, mu = sr.run(mr, mi)
            s_o, p_o = mo.raw_sums(x, q, mu, W[k])
            es = rel(scal[k, 0], s_o)
            ep = np.abs(phase[k] - p_o).max() / np.abs(p_o[0]).max()
            print("table nx=%d elide=%d task %d: scal rel %.2e  phase/max|p11| %.2e   (%.1f ms) stats %s" % (nx, elide, k, es, ep, dt * 1e3, t.last_stats() if k == 0 else ""))
    # determinism
    s2, p2 = t.run(mz, mrel, W, elide=True)
    print("bitwise repeatable:", np.array_equal(s2, scal) and np.array_equal(p2, phase))
    t.close()

# ---- GSF vs oracle
c = np.cos(np.radians(ang))
F = np.stack([0.75 * (1 + c * c), 0.75 * (1 + c * c), 1.5 * c, 1.5 * c, -0.75 * (1 - c * c), 0 * c])
rng = np.random.default_rng(5)
Fs = np.stack([F, F * (1 + 0.3 * np.sin(3 * np.radians(ang))), np.abs(F) + rng.uniform(0, 1, F.shape)])
coef, cn = h.gsf_expand(ang, Fs)
for k in range(3):
    co, cno = go.expand(ang, Fs[k])
    print("gsf cell %d: max abs diff %.3e (max |coef| %.3g), cnorm diff %.2e" % (k, np.abs(coef[k] - co).max(), np.abs(co).max(), abs(cn[k] - cno)))

# ---- bands vs golden
d = np.load(os.path.join(G, "bands.npz"))
for mode in ("GEOS5", "RRTMG", "RRTMGP", "PURDUE"):
    out = h.band_average(d["lam"], d["v"], d[mode + "__lo"], d[mode + "__up"], bool(d[mode + "__usewn"]))
    print("bands %s: rel %.2e" % (mode, rel(out, d[mode + "__avg"])))
print("launches:", h.launch_count())
