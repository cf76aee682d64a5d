import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from geosmie_b200 import _lib
from oracle import mie_oracle as mo
h = _lib.Handle.get(0)
g = np.load(os.path.join(ROOT, "tests/golden/hostlogic.npz"))
ang = np.concatenate([np.linspace(0, 1, 100, endpoint=False), np.linspace(1, 10, 100, endpoint=False), np.linspace(10, 180, 171)])
cost = np.cos(np.radians(ang))
for b in (4, 2):
    x = g["ss__x_%d" % b]
    t0 = time.time(); sr = mo.SizeRange(x, cost); print("bin", b, "nx", x.size, "scipy bessel", time.time() - t0)
    t = _lib.Table(x, sr.nmax, cost, h)
    for (li, rhi) in ((0, 0), (20, 16), (60, 35)):
        key = "ss__cell_%d_%d_%d" % (b, li, rhi)
        mr, mi = g[key + "__m"][:2]
        w = g[key + "__psd"][0]
        t0 = time.time()
        q_o, _, mu_o = sr.run(float(mr), float(mi), nthreads=8)
        s_o, p_o = mo.raw_sums(x, q_o, mu_o, w)
        dt = time.time() - t0
        m = np.array([np.sqrt(complex(mr, mi) ** 2)])
        scal, phase = t.run(m, m, w[None], elide=False)
        q_d, _ = t.particles(m, m, want_s12=False)
        e = np.abs(q_d[0][:, :2] - q_o[:, :2]) / q_o[:, :2]
        nz = w > 0
        print(" cell", (li, rhi), "m=%.4f+%.3ei" % (mr, mi), "oracle %.1fs" % dt, "| per-particle qext/qsca rel err: max %.2e (weighted-nonzero max %.2e), #>1e-9: %d of %d (nonzero-weight: %d)" % (
            e.max(), e[nz].max(), (e.max(axis=1) > 1e-9).sum(), x.size, (e[nz].max(axis=1) > 1e-9).sum()),
            "| bulk scal rel err", np.array2string(np.abs(scal[0, 0] - s_o) / np.maximum(np.abs(s_o), 1e-300), precision=1), "phase/max %.2e" % (np.abs(phase[0] - p_o).max() / np.abs(p_o[0]).max()))
    t.close()
