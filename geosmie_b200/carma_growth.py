"""Sulfate humidification after CARMA (Tabazadeh et al. weight percent, CARMA density table, Kelvin-corrected growth):
the three scalar helpers of src/geosmie/carma_utils.py that the table path needs when a species JSON says
`"rhDep": {"type": "su", "params": {"temp": T}}` (particleparams.humidityGrowth, particleparams.py:101-108).
Host-side scalar code; the rest of carma_utils (bin generators, JSON printers, plotting) is outside the Mie hot path.
"""
import numpy as np

# Tabazadeh fit coefficients (a, b, c, d) at 190 K and 260 K for the three water-activity ranges (carma_utils.py:148-174)
_TABAZ = {
    "low": ((12.37208932, -0.16125516114, -30.490657554, -2.1133114241),
            (13.455394705, -0.1921312255, -34.285174607, -1.7620073078)),
    "mid": ((11.820654354, -0.20786404244, -4.807306373, -5.1727540348),
            (12.891938068, -0.23233847708, -6.4261237757, -4.9005471319)),
    "high": ((-180.06541028, -0.38601102592, -93.317846778, 273.88132245),
             (-176.95814097, -0.36257048154, -90.469744201, 267.45509988)),
}

# CARMA sulfate density table: weight percent nodes, rho = c0 + c1 * T [g cm-3] (carma_utils.py:200-224)
_DN_WTP = np.array([0., 1., 5., 10., 20., 25., 30., 35., 40., 41., 45., 50., 53., 55., 56., 60., 65., 66., 70., 72., 73., 74.,
                    75., 76., 78., 79., 80., 81., 82., 83., 84., 85., 86., 87., 88., 89., 90., 91., 92., 93., 94., 95., 96., 97.,
                    98., 100.])
_DN_C0 = np.array([1., 1.13185, 1.17171, 1.22164, 1.3219, 1.37209, 1.42185, 1.4705, 1.51767, 1.52731, 1.56584, 1.61834, 1.65191,
                   1.6752, 1.68708, 1.7356, 1.7997, 1.81271, 1.86696, 1.89491, 1.9092, 1.92395, 1.93904, 1.95438, 1.98574,
                   2.00151, 2.01703, 2.03234, 2.04716, 2.06082, 2.07363, 2.08461, 2.09386, 2.10143, 2.10764, 2.11283, 2.11671,
                   2.11938, 2.12125, 2.1219, 2.12723, 2.12654, 2.12621, 2.12561, 2.12494, 2.12093])
_DN_C1 = np.array([0., -0.000435022, -0.000479481, -0.000531558, -0.000622448, -0.000660866, -0.000693492, -0.000718251,
                   -0.000732869, -0.000735755, -0.000744294, -0.000761493, -0.000774238, -0.00078392, -0.000788939, -0.00080946,
                   -0.000839848, -0.000845825, -0.000874337, -0.000890074, -0.00089873, -0.000908778, -0.000920012, -0.000932184,
                   -0.000959514, -0.000974043, -0.000988264, -0.00100258, -0.00101634, -0.00102762, -0.00103757, -0.00104337,
                   -0.00104563, -0.00104458, -0.00104144, -0.00103719, -0.00103089, -0.00102262, -0.00101355, -0.00100249,
                   -0.00100934, -0.000998299, -0.000990961, -0.000985845, -0.000984529, -0.000989315])


def wtpct(relhum, temp=220.):
    """Weight percent H2SO4 of a sulfuric-acid droplet at water activity `relhum` (carma_utils.py:136-186)."""
    activ = relhum
    if activ < 0.05:
        activ = np.max([activ, 1.e-6])
        (a1, b1, c1, d1), (a2, b2, c2, d2) = _TABAZ["low"]
    elif (activ >= 0.05) & (activ <= 0.85):
        (a1, b1, c1, d1), (a2, b2, c2, d2) = _TABAZ["mid"]
    else:
        activ = np.min([activ, 1.])
        (a1, b1, c1, d1), (a2, b2, c2, d2) = _TABAZ["high"]
    contl = a1 * (activ ** b1) + c1 * activ + d1
    conth = a2 * (activ ** b2) + c2 * activ + d2
    contt = contl + (conth - contl) * ((temp - 190.) / 70.)
    conwtp = (contt * 98.) + 1000.
    w = (100. * contt * 98.) / conwtp
    return np.min([np.max([w, 1.]), 100.])


def dens(relhum, temp=220.):
    """Density [g cm-3] of the droplet from the CARMA table, linear in weight percent (carma_utils.py:190-241)."""
    wtp = wtpct(relhum, temp=temp)
    i = 0
    while wtp > _DN_WTP[i]:
        i += 1
    den2 = _DN_C0[i] + _DN_C1[i] * temp
    if (i == 0) | (wtp == _DN_WTP[i]):
        return den2
    den1 = _DN_C0[i - 1] + _DN_C1[i - 1] * temp
    frac = (_DN_WTP[i] - wtp) / (_DN_WTP[i] - _DN_WTP[i - 1])
    return den1 * frac + den2 * (1.0 - frac)


def grow_v75(relhum, rd, temp=220.):
    """Wet/dry radius ratio of a sulfate particle of dry radius rd [m] with the Kelvin correction (carma_utils.py:245-314)."""
    rdry = rd * 100.                      # cm
    rhopdry, mw_h2so4, rgas = 1.923, 98., 8.31447e7
    # saturation vapour pressure (Curry & Webster 4.31) and water mass concentration [g cm-3]
    es = 611. * np.exp(2.501e6 / 461. * (1. / 273.16 - 1 / temp))
    n_v = relhum * es / (1.38e-23 * temp)
    navogad = 6.022e23
    h2o_mass = n_v / navogad * 0.018 * 1000. / 1.e6
    # Kelvin effect evaluated at 80 wt %
    wtpkelv = 80.
    den1 = 2.00151 - 0.000974043 * temp
    den2 = 2.01703 - 0.000988264 * temp
    drho_dwt = den2 - den1
    sig1 = 79.3556 - 0.0267212 * temp
    sig2 = 75.608 - 0.0269204 * temp
    dsigma_dwt = (sig2 - sig1) / (85.9195 - 79.432)
    sigkelv = sig1 + dsigma_dwt * (80.0 - 79.432)
    rwet = rdry * (100. * rhopdry / wtpkelv / den2) ** (1. / 3.)
    kb = 1. + wtpkelv * drho_dwt / den2 - 3. * wtpkelv * dsigma_dwt / (2. * sigkelv)
    ka = 2. * mw_h2so4 * sigkelv / (den1 * rgas * temp * rwet)
    h2o_kelv = h2o_mass / np.exp(ka * kb)
    relhum_ = h2o_kelv * navogad / 18. * 1.3807e-16 * temp / (es * 10.)
    rhopwet = dens(relhum_, temp=temp)
    rwet = rdry * (100. * rhopdry / wtpct(relhum_) / rhopwet) ** (1. / 3.)
    return rwet / rdry
