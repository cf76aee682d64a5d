import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from scipy.special import jv, yv
from geosmie_b200 import _lib
from oracle import mie_oracle as mo
h = _lib.Handle.get(0)
ang = np.concatenate([np.linspace(0, 1, 100, endpoint=False), np.linspace(1, 10, 100, endpoint=False), np.linspace(10, 180, 171)])
cost = np.cos(np.radians(ang))
x = np.geomspace(300.0, 2513.0, 64)
sr = mo.SizeRange(x, cost)
for (mr, mi) in ((1.5, 1e-8), (1.5, 1e-3), (1.5, 0.1)):
    q_o, s_o, _ = sr.run(mr, mi, want_s12=True, want_mueller=False)
    t = _lib.Table(x, sr.nmax, cost, h)
    m = np.array([np.sqrt(complex(mr, mi) ** 2)])
    q_d, s_d = t.particles(m, m)
    jh = np.concatenate([jv(np.arange(n + 1) + 0.5, xx) for xx, n in zip(x, sr.nmax)])
    yh = np.concatenate([yv(np.arange(n + 1) + 0.5, xx) for xx, n in zip(x, sr.nmax)])
    t.set_bessel(jh, yh)
    q_s, s_s = t.particles(m, m)
    e_d = np.abs(q_d[0][:, :2] - q_o[:, :2]) / q_o[:, :2]
    e_s = np.abs(q_s[0][:, :2] - q_o[:, :2]) / q_o[:, :2]
    print("m=%g+%gi  device-bessel vs oracle: max %.2e median %.2e | scipy-bessel vs oracle: max %.2e | s12 dev %.2e scipy %.2e" % (
        mr, mi, e_d.max(), np.median(e_d), e_s.max(),
        (np.abs(s_d[0] - s_o).max(axis=(1, 2)) / np.abs(s_o).max(axis=(1, 2))).max(),
        (np.abs(s_s[0] - s_o).max(axis=(1, 2)) / np.abs(s_o).max(axis=(1, 2))).max()))
    t.close()
