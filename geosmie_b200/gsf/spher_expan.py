"""Drop-in for the executable ./spher_expan.x that src/gsf/convertncdf.py spawns per cell (convertncdf.py:186-189):

    python -m geosmie_b200.gsf.spher_expan <file> [<file> ...]

<file>: text, one row per scattering angle, 7 columns `angle[deg] F11 F22 F33 F44 F12 F34` (np.savetxt of allvals.T).
Writes, like main of src/gsf/spher_expan.f (:84-107),
    <file>.expan_coeff   line 1 `L1MAX-1, CNORM`, then `l AL1 AL2 AL3 AL4 BET1 BET2` in '(X,I5,6F17.10)'
    <file>.expan_matr    the matrix re-synthesised from the coefficients, '(F6.2,X,4E15.5,2F11.5)'
and prints `NG fiterr` and `FINAL NG ... Error ...` (:63-66, :83).  The arithmetic runs on the GPU (gm_gsf_diagnose): all
files given on the command line are expanded in ONE call when they share the angle grid, instead of one process per cell.
NSPHER = 129 > 0 (params.h:14), so the adaptive NG loop of the Fortran main (:67-82) never runs in the reference build; it is
available here as `--nspher 0` (what a rebuild with NSPHER <= 0 would do): NG starts at MIN_NSPHER = 129 and grows by DELTA_NG = 10
until the fit error is <= DESIRED_ERR = 0.02 (params.h:9-15), every file with its own final NG.
"""
import sys

import numpy as np

NSPHER = 129      # params.h:14
MIN_NSPHER, DELTA_NG, DESIRED_ERR = 129, 10, 0.02      # params.h:15, :12, :11
NG_LIMIT = 2048   # largest number of Gauss nodes of the device kernels (the Fortran arrays allow NG_MAX = 100000)


def fortran_f(v, w, d):
    """Fortran Fw.d edit descriptor."""
    s = "%*.*f" % (w, d, v)
    if len(s) > w and s.lstrip().startswith(("0.", "-0.")):
        s = s.replace("0.", ".", 1)            # Fortran drops the optional leading zero when the field is too narrow
    return s if len(s) <= w else "*" * w


def fortran_e(v, w, d):
    """Fortran Ew.d edit descriptor as gfortran prints it: 0.ddddd E+xx (two exponent digits, 'E' dropped for three)."""
    if v == 0.0 or not np.isfinite(v):
        body = ("0." + "0" * d + "E+00") if v == 0.0 else ("NaN" if np.isnan(v) else "Infinity")
        sign = "-" if (v == 0.0 and np.signbit(v)) or v == -np.inf else ""
    else:
        m, e = ("%.*e" % (d - 1, abs(v))).split("e")      # D.DDDD, correctly rounded to d significant digits
        digits = m.replace(".", "")
        ex = int(e) + 1
        body = "0." + digits + ("E%+03d" % ex if abs(ex) < 100 else "%+04d" % ex)
        sign = "-" if v < 0 else ""
    s = sign + body
    if len(s) > w and body.startswith("0."):
        s = sign + body[1:]
    return s.rjust(w) if len(s) <= w else "*" * w


def format_expan_coeff(coef, cnorm):
    """coef [6][ng] (normalised AL1,AL2,AL3,AL4,BET1,BET2), cnorm -> text of <file>.expan_coeff (spher_expan.f:93-107)."""
    ng = coef.shape[1]
    lines = [" %5d%s" % (ng - 1, fortran_f(cnorm, 17, 10))]
    for l in range(ng):
        lines.append(" %5d%s" % (l, "".join(fortran_f(coef[k, l], 17, 10) for k in range(6))))
    return "\n".join(lines) + "\n"


def format_expan_matr(ang_deg, fout):
    """fout [6][nang] -> text of <file>.expan_matr (spher_expan.f:84-90).  The Fortran prints angl(i)*R2D after the D2R scaling."""
    pi = np.arccos(-1.0)
    back = (np.asarray(ang_deg, dtype=float) * (pi / 180.0)) * (180.0 / pi)
    lines = []
    for i in range(fout.shape[1]):
        lines.append(fortran_f(back[i], 6, 2) + " " + "".join(fortran_e(fout[k, i], 15, 5) for k in range(4)) +
                     "".join(fortran_f(fout[k, i], 11, 5) for k in (4, 5)))
    return "\n".join(lines) + "\n"


def expand_files(paths, handle=None, ng=NSPHER):
    """Expand every input file; files sharing an angle grid go to the GPU together.  Returns {path: fiterr}."""
    from .. import _lib
    h = handle or _lib.Handle.get()
    data = [np.loadtxt(p, ndmin=2) for p in paths]
    for p, a in zip(paths, data):
        if a.shape[1] != 7:
            raise ValueError("%s: expected 7 columns (angle F11 F22 F33 F44 F12 F34), found %d" % (p, a.shape[1]))
        if a.shape[0] >= 1000:
            raise ValueError("READMATRIX: Too many angles in input %s" % p)      # NANG_MAX, params.h:1, spher_expan.f:229-232
    out = {}
    done = [False] * len(paths)
    for i in range(len(paths)):
        if done[i]:
            continue
        same = [j for j in range(i, len(paths)) if not done[j] and data[j].shape == data[i].shape and
                np.array_equal(data[j][:, 0], data[i][:, 0])]
        ang = data[i][:, 0]
        F = np.stack([data[j][:, 1:].T for j in same])
        if ng > 0:
            coef, cn, fout, fiterr = h.gsf_diagnose(ang, F, ng)
            res = [(ng, [(ng, float(fiterr[k]))], coef[k], cn[k], fout[k], float(fiterr[k])) for k in range(len(same))]
        else:
            res = expand_adaptive(h, ang, F)
        for (ngk, trail, coefk, cnk, foutk, errk), j in zip(res, same):
            done[j] = True
            print(paths[j])
            for n_, e_ in trail:
                print(" %11d  %s" % (n_, repr(e_)))
            print(" FINAL NG %11d Error  %s" % (ngk, repr(errk)))
            with open(paths[j] + ".expan_matr", "w") as f:
                f.write(format_expan_matr(ang, foutk))
            with open(paths[j] + ".expan_coeff", "w") as f:
                f.write(format_expan_coeff(coefk, cnk))
            out[paths[j]] = errk
    return out


def expand_adaptive(h, ang, F, min_ng=MIN_NSPHER, delta=DELTA_NG, desired=DESIRED_ERR, limit=NG_LIMIT):
    """The NSPHER <= 0 branch of the Fortran main (spher_expan.f:67-82) for a batch of matrices on one angle grid: every matrix is
    expanded with NG = min_ng, min_ng + delta, ... until its fit error (one_calc) is <= desired; matrices still above the threshold go
    to the GPU together at the next NG.  Returns per matrix (final NG, [(NG, fiterr), ...], coef [6][NG], cnorm, fout [6][nang], fiterr)."""
    n = F.shape[0]
    res = [None] * n
    trail = [[] for _ in range(n)]
    pending = list(range(n))
    ng = min_ng
    while pending:
        coef, cn, fout, fiterr = h.gsf_diagnose(ang, F[pending], ng)
        still = []
        for k, j in enumerate(pending):
            e = float(fiterr[k])
            trail[j].append((ng, e))
            if e <= desired or ng + delta > limit:
                res[j] = (ng, trail[j], coef[k], float(cn[k]), fout[k], e)
            else:
                still.append(j)
        pending = still
        ng += delta
    return res


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    ng = NSPHER
    if "--nspher" in argv:                      # the reference fixes NSPHER at compile time (params.h:14); <= 0 selects the adaptive loop
        k = argv.index("--nspher")
        ng = int(argv[k + 1])
        argv = argv[:k] + argv[k + 2:]
    if len(argv) < 1:
        print(" Wrong number of arguments: %d" % len(argv))
        print(" Usage: spher_expan [--nspher N] <filename> [<filename> ...]")
        return 1
    expand_files(argv, ng=ng)
    return 0


if __name__ == "__main__":
    sys.exit(main())
