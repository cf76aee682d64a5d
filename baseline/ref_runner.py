"""Times the UNMODIFIED reference (GEOS-ESM/GEOSmie, Python + numba) on the host cores: the CPU arm of bench.py.

The reference tree is imported from baseline/_ref (installed by baseline/install_reference.sh; git-ignored, travels to the GPU box
with the snapshot) or, in the build container, from /root/reference.  Nothing of geosmie_b200's compute path is used here.

What is timed is the reference's own per-cell work of dointegration.fun (:844-889): getHumidRefractiveIndex + calculatePSD + rawMie
(MultipleMie.calculateS12SizeRange: single_mie_coeff_numba + runS12Loop + mie_props_raw per particle, then the list -> array
repacking and calculateScatVals) + integratePSD.  The per-bin pre-computation (MultipleMie.preCalculate: scipy Bessel tables and
the pi/tau tables per distinct nmax) sits outside the reference's cell loops and is done once per bin here as well, untimed; so is
numba's JIT compilation (one warm-up cell).

Sampling (a whole optics_SS table takes ~11 h of one core): a sample cell is evaluated on every `stride`-th point of the bin's size
grid (the cost per particle depends on x only, so a regular sub-grid has the table's cost mix), and a "cell set" holds one cell of
EVERY bin of the species -- every bin has the same number of cells (61 x 36) and nearly the same number of grid points, so cell
sets sample the table proportionally and  evals / time  of a set is the table-equivalent rate.  All host cores are used the way the
reference's own production script does it (src/scripts/proc.v2.1.0.csh:30-33: one process per table): a fork pool with one worker
per core, every worker evaluating whole cell sets.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_S = {}


def reference_root():
    for r in (os.path.join(HERE, "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(r, "src", "pymiecoated", "pymiecoated")):
            return r
    return None


def _setup(sp):
    """Imports the reference (once) and prepares the species' per-bin inputs.  Returns the state dict."""
    if sp in _S:
        return _S[sp]
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (baseline/_ref missing: run baseline/install_reference.sh where /root/reference exists)")
    os.environ["GEOSMIE_REFERENCE"] = root
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("NUMBA_NUM_THREADS", "1")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refharness as rh
    rh.REF = root
    R = rh.reference()
    from scipy.interpolate import interp1d
    with rh.reference_cwd():
        params = R.particleparams.getParticleParams("geosparticles/%s.json" % sp, "json")
        water = R.particleparams.getWaterM()
    ang = np.concatenate([np.linspace(0., 1., 100, endpoint=False), np.linspace(1., 10., 100, endpoint=False),
                          np.linspace(10., 180., 171, endpoint=True)])
    ml = params["mList"]
    rh_used = np.array(params["rh"])
    if "maxrh" in params:
        rh_used[np.where(rh_used > params["maxrh"])[0]] = params["maxrh"]
    pp_ = params["psd"]["params"]
    nbin = len(pp_["r0"] if params["psd"]["type"] == "lognorm" else pp_["rMinMaj"])
    _S[sp] = dict(R=R, params=params, water=water, cost=np.cos(np.radians(ang)), interp1d=interp1d, ml=ml, rh=rh_used, nbin=nbin,
                  mm={}, root=root)
    return _S[sp]


def prepare(sp, stride):
    """Per-bin pre-computation on the strided sub-grids (untimed, like MultipleMie.preCalculate in fun :797-798) + JIT warm-up."""
    s = _setup(sp)
    DI = s["R"].dointegration
    lam_all = s["ml"][0][0]
    t0 = time.time()
    for b in range(s["nbin"]):
        if (b, stride) in s["mm"]:
            continue
        xx, dr = DI.initializeXarr(s["params"], b, lam_all[0], lam_all[-1])
        sel = np.arange(0, xx.size, stride)
        mm = s["R"].pymiecoated_mie_coated.MultipleMie(xx[sel], None, s["cost"])
        mm.preCalculate()
        s["mm"][(b, stride)] = (mm, xx, dr, sel)
    cell(sp, 0, stride, 0, 0)          # numba compilation
    return time.time() - t0, [int(s["mm"][(b, stride)][1].size) for b in range(s["nbin"])]


def cell(sp, b, stride, li, rhi):
    """One (wavelength, RH) cell of bin b with the reference's statements (dointegration.py:813-889).  Returns particle-evals."""
    s = _setup(sp)
    DI, params, water, interp1d, ml, rh_used = s["R"].dointegration, s["params"], s["water"], s["interp1d"], s["ml"], s["rh"]
    mm, xx, dr, sel = s["mm"][(b, stride)]
    lam = ml[0][0][li]
    nref0 = [complex(interp1d(ml[i][0], ml[i][1])(lam), -interp1d(ml[i][0], ml[i][2])(lam)) for i in range(len(ml))]
    nw = complex(interp1d(water[0], water[1])(lam), interp1d(water[0], water[2])(lam))
    _, _, _, rrat0 = DI.getHumidRefractiveIndex(params, b, 0, rh_used, nref0, nw)
    _, reff_mass0, _, _ = DI.calculatePSD(params, b, 0., rh_used, xx, dr, rrat0, lam)
    mr, mi, gf, rrat = DI.getHumidRefractiveIndex(params, b, rhi, rh_used, nref0, nw)
    psd, ref, rlow, rup = DI.calculatePSD(params, b, rh_used[rhi], rh_used, xx, dr, rrat, lam)
    rhop0 = params["rhop0"][b] if isinstance(params["rhop0"], list) else params["rhop0"]
    rhop = rrat ** 3. * rhop0 + (1. - rrat ** 3.) * 1000.
    allret = [DI.rawMie(mm, DI.scatkeys, DI.scalarkeys, lam, mr[i], mi[i], None, s["cost"]) for i in range(len(mr))]
    if len(allret) == 1:
        allret = [allret[0] for _ in range(len(psd))]
    psd_sub = [np.asarray(p)[sel] for p in psd]
    with np.errstate(all="ignore"):      # a strided sub-grid may miss every populated point of a narrow bin (0/0 in the ratios)
        DI.integratePSD(mm.xArr, allret, psd_sub, params["psd"]["params"]["fracs"][b], lam, reff_mass0, rhop0, rhop)
    return int(sel.size) * len(mr)


def cell_set(args):
    """One cell of every bin (a proportional sample of the table).  Returns (particle-evals, seconds)."""
    sp, stride, li, rhi = args
    s = _setup(sp)
    t0 = time.perf_counter()
    n = 0
    for b in range(s["nbin"]):
        n += cell(sp, b, stride, li, rhi)
    return n, time.perf_counter() - t0


class Farm(object):
    """Fork pool with one worker per core; the parent has imported the reference, built the per-bin tables and compiled the numba
    functions before the fork, so the workers inherit all of it."""

    def __init__(self, sp, stride, cores=None):
        import multiprocessing as mp
        self.sp, self.stride = sp, stride
        self.cores = cores or len(os.sched_getaffinity(0))
        self.prep_s, self.nx = prepare(sp, stride)
        self.pool = mp.get_context("fork").Pool(self.cores) if self.cores > 1 else None
        self.k = 0

    def step(self, sets_per_core=1):
        """Evaluates cores x sets_per_core cell sets (different cells every call).  Returns (evals, wall seconds, core seconds)."""
        n = self.cores * sets_per_core
        items = []
        for _ in range(n):
            li, rhi = (7 * self.k) % 61, (5 * self.k) % 36       # walks over the whole (wavelength, RH) grid
            items.append((self.sp, self.stride, li, rhi))
            self.k += 1
        t0 = time.perf_counter()
        res = self.pool.map(cell_set, items, chunksize=sets_per_core) if self.pool else [cell_set(it) for it in items]
        wall = time.perf_counter() - t0
        return sum(r[0] for r in res), wall, sum(r[1] for r in res)

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()
            self.pool = None


def describe(sp, stride, cores, nsets, nx):
    return ("unmodified reference (rawMie + integratePSD per cell, numba) on %d cores: %d cell sets = one (lambda, RH) cell of each of the "
            "%d bin(s) of %s.json, every %d-th point of the %s-point size grids, 371 angles; per-bin preCalculate and JIT untimed"
            % (cores, nsets, len(nx), sp, stride, "/".join(str(n) for n in nx)))


if __name__ == "__main__":
    sp = sys.argv[1] if len(sys.argv) > 1 else "ss"
    stride = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    f = Farm(sp, stride, cores=int(sys.argv[3]) if len(sys.argv) > 3 else None)
    print("prepared in %.1f s" % f.prep_s, f.nx)
    for _ in range(3):
        n, wall, core = f.step()
        print("evals %d wall %.2f s core %.2f s -> %.0f evals/s (%d cores), %.0f evals/s/core" % (n, wall, core, n / wall, f.cores, n / core))
    f.close()
