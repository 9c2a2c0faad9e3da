"""
model_host.py -- numpy mirror of the Goloskokov-Kroll model GPD H(x, xi, t) (reference:
tiktaalik/model.py:22-114 -> f90wrap/model.f90:16-158 -> f90src/model/gk.f90:75-241) for machines without
a GPU (building inputs, plotting).  The package's model API (tiktaalik_b200.model) runs on the device;
this module is reachable as tiktaalik_b200.model.host.  Same function names, argument conventions and
output shape (nx, nxi, nt) as the reference; the model is an input generator, not part of the hot path.
"""

import numpy as np

# gk.f90:30-56
_a0_val, _ap_val, _a0_sea, _ap_sea, _kappa_s = 0.48, 0.9, 1.1, 0.15, 1.68
# gk.f90:59-64, C(i, j): i = 0 gluon, 1 u_val, 2 d_val, 3 strange
_C = np.array([[2.23, 5.43, -34.0, 40.6],
               [1.52, 2.88, -0.095, 0.0],
               [0.76, 3.11, -3.99, 0.0],
               [0.123, -0.327, 0.692, -0.486]])
_delta = np.array([_a0_sea - 1.0, _a0_val, _a0_val, _a0_sea])
_ap = np.array([_ap_sea, _ap_val, _ap_val, _ap_sea])


def _grid(x, xi, t):
    x, xi, t = (np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in (x, xi, t))
    return x[:, None, None], xi[None, :, None], t[None, None, :]


def _kw(x, xi, w):                     # gk.f90:193-209
    if w == 1:
        return 1.0 - x**2 / xi**2
    if w == 2:
        return 2.0 * (x / xi**2 - 1.0)
    return 1.0 - 1.0 / xi**2 + 0.0 * x


def _ew(x, xi, w):                     # gk.f90:211-231
    k1, k2, k3 = _kw(x, xi, 1), _kw(x, xi, 2), _kw(x, xi, 3)
    return {1: k1**2, 2: 2.0 * k1 * k2, 3: k2**2 + 2.0 * k1 * k3, 4: 2.0 * k2 * k3, 5: k3**2}[w]


def _pw(i, j, w, t):                   # gk.f90:233-239
    return (j / 2.0 + w) - _delta[i] - _ap[i] * t


def _term(x, xi, t, z, i, nw):
    """H_n1_term (nw = 3, gk.f90:161-175) / H_n2_term (nw = 5, gk.f90:177-191)"""
    H = 0.0
    with np.errstate(invalid="ignore", divide="ignore"):
        for j in range(4):
            for w in range(1, nw + 1):
                p = _pw(i, j, w, t)
                coef = _kw(x, xi, w) if nw == 3 else _ew(x, xi, w)
                H = H + _C[i, j] * coef * np.power(z, p) / p
    return H * 0.75 / xi if nw == 3 else H * 1.875 / xi / 2.0


def _H_n1(x, xi, t, i):                # gk.f90:125-139
    x, xi, t = _grid(x, xi, t)
    zi = (x - xi) / (1.0 - xi)
    zf = (x + xi) / (1.0 + xi)
    dglap = x > xi
    erbl = (~dglap) & (x > -xi)
    tf = _term(x, xi, t, np.where(dglap | erbl, zf, 1.0), i, 3)
    ti = _term(x, xi, t, np.where(dglap, zi, 1.0), i, 3)
    return np.where(dglap, tf - ti, np.where(erbl, tf, 0.0))


def _H_n2(x, xi, t, i, sgn):           # gk.f90:141-159
    x, xi, t = _grid(x, xi, t)
    zi = (x - xi) / (1.0 - xi)
    zf = (x + xi) / (1.0 + xi)
    zj = (-x - xi) / (1.0 - xi)
    zm = (-x + xi) / (1.0 + xi)
    dglap = x > xi
    erbl = (~dglap) & (x > -xi)
    neg = ~(dglap | erbl)
    a = _term(x, xi, t, np.where(dglap | erbl, zf, 1.0), i, 5)
    b = _term(x, xi, t, np.where(dglap, zi, 1.0), i, 5)
    c = _term(-x, xi, t, np.where(erbl | neg, zm, 1.0), i, 5)
    d = _term(-x, xi, t, np.where(neg, zj, 1.0), i, 5)
    return np.where(dglap, a - b, np.where(erbl, a + sgn * c, sgn * (c - d)))


# ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
# Model GPDs (reference: model.py:22-114; gk.f90:75-123)

def Hu_val(x, xi, t):
    return _H_n1(x, xi, t, 1)

def Hd_val(x, xi, t):
    return _H_n1(x, xi, t, 2)

def Hu_sea(x, xi, t):
    return _kappa_s * _H_n2(x, xi, t, 3, -1.0)

def Hd_sea(x, xi, t):
    return _kappa_s * _H_n2(x, xi, t, 3, -1.0)

def Hu(x, xi, t):
    return Hu_val(x, xi, t) + Hu_sea(x, xi, t)

def Hd(x, xi, t):
    return Hd_val(x, xi, t) + Hd_sea(x, xi, t)

def Hs(x, xi, t):
    return _H_n2(x, xi, t, 3, -1.0)

def Hg(x, xi, t):
    return _H_n2(x, xi, t, 0, 1.0)

def H_singlet(x, xi, t):
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    f = (
              Hu(x,xi,t) - Hu(-x,xi,t)
            + Hd(x,xi,t) - Hd(-x,xi,t)
            + Hs(x,xi,t) - Hs(-x,xi,t)
            )
    return f
