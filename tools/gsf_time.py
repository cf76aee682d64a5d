#!/usr/bin/env python3
"""Time the GSF expansion of gm_table_run's phase layout for `ncell` cells (device pointers, CUDA events)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from geosmie_b200 import _lib, dointegration as DI
ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 2196
h = _lib.Handle.get(0)
ang = DI.table_angles()
c = np.cos(np.radians(ang))
rng = np.random.default_rng(0)
g = rng.uniform(0.1, 0.9, ncell)[:, None]
hg = (1 - g * g) / (1 + g * g - 2 * g * c[None]) ** 1.5
P4 = torch.from_numpy(np.stack([hg, -0.3 * hg * (1 - c * c), hg * c, 0.1 * hg * np.sin(np.radians(ang))], axis=1)).cuda()
coef = torch.empty((ncell, 6, 129), dtype=torch.float64, device="cuda")
cn = torch.empty(ncell, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream(); h.set_stream(st.cuda_stream)
for _ in range(3):
    h.gsf_expand_phase4_dev(ang, ncell, P4.data_ptr(), coef.data_ptr(), cn.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20):
    h.gsf_expand_phase4_dev(ang, ncell, P4.data_ptr(), coef.data_ptr(), cn.data_ptr())
e1.record(); torch.cuda.synchronize()
print("GSF %d cells: %.4f ms per call (%s)" % (ncell, e0.elapsed_time(e1) / 20, "k_gsf" if os.environ.get("GEOSMIE_GSF_SIMPLE") else "k_gsf_multi"))
