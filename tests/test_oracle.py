"""CPU tests (-m "not gpu"): pin the oracle against the reference's golden vectors and reference-generated fixtures."""
import numpy as np
import pytest

from helpers import load, relerr, table_angles
from oracle import gsf_oracle as go
from oracle import mie_oracle as mo

# tolerance of the reference's own test-suite: epsilon = 1e3 * DBL_EPSILON (test_mie.py:30)
EPS = 1e3 * np.finfo(float).eps

# golden values copied from src/pymiecoated/pymiecoated/test/test_mie.py:45-109
REFERENCE_GOLDENS = [
    (dict(x=2.5, eps=(1.5 + 0.5j) ** 2),
     [2.562873497454734, 1.0970718190883924, 1.4658016783663417, 0.12358646817981821, 0.74890597894850719, 0.112651210275834],
     (-0.49958438416709694 - 0.24032581667666403j, 0.11666852712178288 + 0.051661382367147853j)),
    (dict(x=1.5, y=5.0, eps=(1.5 + 0.5j) ** 2, eps2=(1.2 + 0.2j) ** 2),
     [2.0765452928100769, 0.90777572021757091, 1.168769572592506, 0.022692436240597712, 0.90560220988567752, 0.024997844440209204],
     (0.28677219451960079 - 0.063605895700765691j, -0.32635924647084191 + 0.12670342074119806j)),
    (dict(x=4.0, eps=2.2 + 0.8j, mu=1.6 + 1.4j),
     [2.6665582594291073, 1.1255460946883893, 1.541012164740718, 0.0072453174040961301, 0.89955981937838192, 0.0064371574281033928],
     (0.14683196954000932 - 0.017479181764394575j, -0.12475414168001844 + 0.28120475717321358j)),
]


@pytest.mark.parametrize("kw,q_ref,s_ref", REFERENCE_GOLDENS)
def test_reference_unit_test_goldens(kw, q_ref, s_ref):
    an, bn, nmax, size = mo.mie_coeffs(kw["x"], kw["eps"], kw.get("mu", 1.0), kw.get("y"), kw.get("eps2"))
    q = mo.mie_props(an, bn, nmax, size)
    assert relerr(q, q_ref) < EPS
    s1, s2 = mo.mie_s12(an, bn, nmax, -0.6)
    assert abs(s1 - s_ref[0]) / abs(s_ref[0]) < EPS
    assert abs(s2 - s_ref[1]) / abs(s_ref[1]) < EPS


def test_oracle_vs_reference_single_particles():
    d = load("mie_single.npz")
    par, us, qg, sg = d["par"], d["us"], d["q"], d["s12"]
    for i in range(par.shape[0]):
        x, y = par[i, 0], par[i, 1]
        eps, mu = complex(par[i, 2], par[i, 3]), complex(par[i, 4], par[i, 5])
        coated = not np.isnan(y)
        an, bn, nmax, size = mo.mie_coeffs(x, eps, 1.0 if coated else mu, y if coated else None,
                                           complex(par[i, 6], par[i, 7]) if coated else None)
        q = mo.mie_props(an, bn, nmax, size)
        # qabs = qext - qsca cancels for weak absorbers: compare it on the scale of qext
        scale = np.abs(qg[i]).copy()
        scale[2] = max(scale[2], abs(qg[i][0]))
        # below x ~ 0.1 the reference formula itself amplifies 1-ulp differences (hard part 1)
        assert np.max(np.abs(q - qg[i]) / scale) < (1e-11 if size >= 0.1 else 1e-9), (i, x, y)
        for k, u in enumerate(us):
            s1, s2 = mo.mie_s12(an, bn, nmax, u)
            ref = sg[i, k]
            smax = np.abs(sg[i]).max()
            assert abs(s1 - complex(ref[0], ref[1])) / smax < 1e-12
            assert abs(s2 - complex(ref[2], ref[3])) / smax < 1e-12


def test_oracle_vs_reference_size_range():
    d = load("size_range.npz")
    for ci in range(int(d["ncase"])):
        x = d["x_%d" % ci]
        mr, mi = d["m_%d" % ci]
        sr = mo.SizeRange(x, d["cost"])
        q, s12, mu = sr.run(float(mr), float(mi), want_s12=True, want_mueller=True)
        qg, sg = d["q_%d" % ci], d["s12_%d" % ci]
        scale = np.abs(qg).copy()
        scale[:, 2] = np.maximum(scale[:, 2], np.abs(qg[:, 0]))
        scale[:, 5] = np.maximum(scale[:, 5], 1e-300)
        assert np.max(np.abs(q - qg) / np.maximum(scale, 1e-300)) < 1e-11
        assert np.max(np.abs(s12 - sg)) / np.abs(sg).max() < 1e-12
        # Mueller elements of calculateScatVals (dointegration.py:1044-1050)
        s1 = sg[..., 0] + 1j * sg[..., 1]
        s2 = sg[..., 2] + 1j * sg[..., 3]
        p11 = 0.5 * (np.abs(s1) ** 2 + np.abs(s2) ** 2)
        assert np.max(np.abs(mu[0] - p11)) / p11.max() < 1e-12
        assert np.max(np.abs(mu[4] + (np.conj(s1) * s2).imag)) / p11.max() < 1e-12


def test_oracle_integrate_psd_vs_reference():
    d = load("cells.npz")
    x, lam = d["x"], float(d["lam"])
    cost = np.cos(np.radians(d["ang"]))
    sr = mo.SizeRange(x, cost)
    for ci in range(int(d["ncase"])):
        ms = d["m_%d" % ci]
        psd = list(d["psd_%d" % ci])
        fracs = list(d["fracs_%d" % ci])
        rhop0, rhop = d["rhop_%d" % ci]
        allret = [mo.raw_mie(sr, lam, float(mr), float(mi)) for mr, mi in ms]
        if len(allret) == 1:
            allret = [allret[0] for _ in psd]
        ret = mo.integrate_psd(x, allret, psd, fracs, lam, list(d["reff0_%d" % ci]), rhop0, rhop)
        for k in ("qext", "qsca", "qabs", "qb", "g", "csca", "cext", "bsca", "bext", "bbck", "area", "volume", "mass", "rEff", "rMass"):
            ref = d["ret_%d__%s" % (ci, k)]
            assert relerr(ret[k], ref) < 1e-11, (ci, k)
        for k in ("p11", "p12", "p33", "p34"):
            ref = d["ret_%d__%s" % (ci, k)]
            assert np.max(np.abs(ret[k] - ref)) / np.abs(d["ret_%d__p11" % ci]).max() < 1e-12, (ci, k)


# ---------------------------------------------------------------------------------------------- GSF
def rayleigh(ang):
    c = np.cos(np.radians(ang))
    return np.stack([0.75 * (1 + c * c), 0.75 * (1 + c * c), 1.5 * c, 1.5 * c, -0.75 * (1 - c * c), 0 * c])


def test_gsf_gauss_nodes_match_numpy():
    z, w = go.gauss(129)
    zn, wn = np.polynomial.legendre.leggauss(129)
    assert np.max(np.abs(z - zn)) < 1e-14
    assert np.max(np.abs(w - wn)) < 1e-14      # IND1 = 0: weights on [-1, 1]


def test_gsf_rayleigh_known_answer():
    """Analytic Rayleigh expansion: a1 = (1,0,1/2), a2 = (0,0,3), a4 = (0,3/2,0), b1 = (0,0,sqrt(6)/2), rest 0.
    On the 371-angle grid the linear interpolation of spher_expan.f limits the agreement to ~1e-4;
    a 0.02-degree grid (below NANG_MAX = 1000 is impossible, so 0.2 degrees) tightens it."""
    for ang, tol in ((table_angles(), 2e-4), (np.linspace(0, 180, 901), 2e-5)):
        coef, cn = go.expand(ang, rayleigh(ang))
        exact = np.zeros((6, 129))
        exact[0, 0], exact[0, 2] = 1.0, 0.5
        exact[1, 2] = 3.0
        exact[3, 1] = 1.5
        exact[4, 2] = np.sqrt(6.0) / 2.0
        assert np.max(np.abs(coef - exact)) < tol
        assert abs(cn - 1.0) < tol


def test_gsf_matr_round_trip():
    """Expansion followed by re-synthesis with the MATR recurrences (spher_expan.f:419-517) returns the input."""
    ang = table_angles()
    F = rayleigh(ang)
    g = 0.6   # add a smooth forward-peaked Henyey-Greenstein-like F11 so that many orders are exercised
    c = np.cos(np.radians(ang))
    hg = (1 - g * g) / (1 + g * g - 2 * g * c) ** 1.5
    F2 = F * hg
    coef, cn = go.expand(ang, F2)
    back = go.matr(coef / cn, ang)     # undo the CNORM normalisation
    assert np.max(np.abs(back - F2)) / np.abs(F2).max() < 5e-4
    q = go.expand(ang, F2, quantize10=True)[0]
    assert np.max(np.abs(q - coef)) <= 0.5e-10 + 1e-15


def test_gsf_one_calc_diagnostics():
    """one_calc's fit error (spher_expan.f:168-177): the alternative (mid-point) angle grid of READMATRIX :237-258, MATR on the
    un-normalised coefficients and ERREVAL (MAXABS).  Rayleigh is reproduced to the linear-interpolation limit of the grid;
    a matrix that 129 terms cannot represent (a step) must give a large error."""
    ang = table_angles()
    F = rayleigh(ang)
    fout, err = go.one_calc(ang, F)
    coef, cn = go.expand(ang, F)
    assert np.max(np.abs(fout - go.matr(coef / cn, ang))) < 1e-12           # same re-synthesis up to the CNORM round trip
    assert np.max(np.abs(fout[0] - F[0])) <= err < 1e-4
    # the alt grid only matters when the error between the nodes is larger than at the nodes
    uni = np.linspace(0., 180., 181)
    Fu = rayleigh(uni)
    _, err_u = go.one_calc(uni, Fu)
    assert err_u < 2e-4
    step = Fu.copy()
    step[0, 60:] += 1.0
    _, err_s = go.one_calc(uni, step)
    assert err_s > 0.05


def test_fortran_edit_descriptors():
    """Formats of <file>.expan_coeff '(X,I5,6F17.10)' and <file>.expan_matr '(F6.2,X,4E15.5,2F11.5)', spher_expan.f:84-107."""
    from geosmie_b200.gsf import spher_expan as se
    assert se.fortran_e(1.0, 15, 5) == "    0.10000E+01"
    assert se.fortran_e(-0.00123456, 15, 5) == "   -0.12346E-02"
    assert se.fortran_e(0.0, 15, 5) == "    0.00000E+00"
    assert se.fortran_e(9.99999e4, 15, 5) == "    0.10000E+06"      # rounding carries into the exponent
    assert se.fortran_e(1.5e-120, 15, 5) == "    0.15000-119"       # three-digit exponents drop the 'E'
    assert se.fortran_f(180.0, 6, 2) == "180.00" and se.fortran_f(0.5, 6, 2) == "  0.50"
    assert se.fortran_f(-0.75, 11, 5) == "   -0.75000" and se.fortran_f(1e7, 6, 2) == "******"
    coef = np.arange(12, dtype=float).reshape(6, 2) / 7.0
    txt = se.format_expan_coeff(coef, 0.999987)
    rows = txt.splitlines()
    assert rows[0] == "     1     0.9999870000" and len(rows) == 3 and len(rows[1]) == 6 + 6 * 17
    # the reference's consumer: np.loadtxt(skiprows=1, unpack=True), columns 1..6 (convertncdf.py:189, :381-394)
    import io
    back = np.loadtxt(io.StringIO(txt), skiprows=1, unpack=True)
    assert back.shape == (7, 2) and np.max(np.abs(back[1:] - coef)) <= 0.5e-10
    m = se.format_expan_matr(np.array([0.0, 90.0, 180.0]), np.ones((6, 3)))
    assert m.splitlines()[2] == "180.00" + " " + "    0.10000E+01" * 4 + "    1.00000" * 2


def test_spher_expan_text_protocol_with_oracle_backend(tmp_path):
    """The drop-in for ./spher_expan.x (file grouping, formats, what the reference's consumer reads back) with the oracle
    standing in for the GPU handle: input text -> <file>.expan_coeff / <file>.expan_matr (convertncdf.py:185-189, :381-394)."""
    from geosmie_b200.gsf import spher_expan as se

    class OracleHandle(object):
        calls = 0

        def gsf_diagnose(self, ang, F, ng=129, quantize10=False):
            OracleHandle.calls += 1
            co, cn, fo, er = [], [], [], []
            for k in range(F.shape[0]):
                c, n = go.expand(ang, F[k], ng)
                f, e = go.one_calc(ang, F[k], ng)
                co.append(c); cn.append(n); fo.append(f); er.append(e)
            return np.array(co), np.array(cn), np.array(fo), np.array(er)

    ang = table_angles()
    uni = np.linspace(0., 180., 181)
    files, mats = [], []
    for i, a in enumerate((ang, uni, ang)):            # two angle grids: files 0 and 2 are expanded together
        F = rayleigh(a) * (1.0 + 0.1 * i)
        allvals = np.zeros((7, a.size))
        allvals[0], allvals[1:] = a, F
        fn = str(tmp_path / ("x.tempfile%d.txt" % i))
        np.savetxt(fn, allvals.T)
        files.append(fn)
        mats.append((a, F))
    out = se.expand_files(files, handle=OracleHandle())
    assert OracleHandle.calls == 2 and set(out) == set(files)
    for fn, (a, F) in zip(files, mats):
        newdata = np.loadtxt(fn + ".expan_coeff", skiprows=1, unpack=True)
        co, cn = go.expand(a, F, quantize10=True)
        assert newdata.shape == (7, 129) and np.max(np.abs(newdata[1:] - co)) <= 1e-10 + 1e-15
        assert abs(newdata[1, 0] - 1.0) < 1e-9                      # normalised by AL1(0)
        matr = np.loadtxt(fn + ".expan_matr")
        assert matr.shape == (a.size, 7) and np.max(np.abs(matr[:, 0] - a)) <= 0.005
        assert np.max(np.abs(matr[:, 1] - F[0])) < 2e-4 * 1.2        # re-synthesis reproduces F11 to the interpolation limit
        assert out[fn] < 2.5e-4
    bad = str(tmp_path / "bad.txt")
    np.savetxt(bad, np.zeros((5, 3)))
    with pytest.raises(ValueError):
        se.expand_files([bad], handle=OracleHandle())
