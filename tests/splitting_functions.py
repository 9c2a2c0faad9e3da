"""Literature known answers for the forward limit of the evolution kernels: the LO and NLO (MS-bar) DGLAP
splitting functions at x < 1, written with harmonic polylogarithms up to weight 2.

Unpolarised NLO: S. Moch, J. Vermaseren, A. Vogt, Nucl. Phys. B688 (2004) 101, eqs. (4.5)-(4.6) (non-singlet) and
A. Vogt, S. Moch, J. Vermaseren, Nucl. Phys. B691 (2004) 129, eqs. (4.7)-(4.10) (singlet); originally
Curci-Furmanski-Petronzio (1980), Furmanski-Petronzio (1980).  Normalisation there: P = sum a_s^(n+1) P^(n),
a_s = alpha_s / (4 pi).  The polarised LO functions and the polarised NLO pure-singlet term are from
W. Vogelsang, Nucl. Phys. B475 (1996) 47.

None of this comes from the reference under test; the GPD kernels must reduce to these functions as xi -> 0
(Belitsky-Freund-Mueller, the paper the reference implements, sec. 5).
"""
import numpy as np
import mpmath as mp

CF, CA, TF = 4.0 / 3.0, 3.0, 0.5
Z2 = np.pi ** 2 / 6


def _li2(x):
    return float(mp.polylog(2, x))


def _H(x):
    lx, l1 = np.log(x), np.log(1 - x)
    return {"0": lx, "00": 0.5 * lx ** 2, "1": -l1, "2": _li2(x), "10": -lx * l1 - _li2(x), "11": 0.5 * l1 ** 2,
            "m10": lx * np.log(1 + x) + _li2(-x)}


def pqq(x):
    return 2 / (1 - x) - 1 - x


def pqg(x):
    return x * x + (1 - x) ** 2


def pgq(x):
    return (1 + (1 - x) ** 2) / x


def pgg(x):
    return 1 / (1 - x) + 1 / x - 2 + x - x * x


# ---- LO, x < 1, a_s normalisation ----
def P0_qq(x):
    return CF * pqq(x)


def P0_qg(x, nf):
    return 2 * nf * pqg(x)


def P0_gq(x):
    return 2 * CF * pgq(x)


def P0_gg(x):
    return 4 * CA * pgg(x)


def dP0_qg(x, nf):
    return 2 * nf * (2 * x - 1)


def dP0_gq(x):
    return 2 * CF * (2 - x)


def dP0_gg(x):
    return 4 * CA * (1 / (1 - x) - 2 * x + 1)


# ---- NLO unpolarised, x < 1 ----
def P1_ns(x, sign, nf):
    h = _H(x)
    plus = (4 * CA * CF * (pqq(x) * (67 / 18 - Z2 + 11 / 6 * h["0"] + h["00"]) + pqq(-x) * (Z2 + 2 * h["m10"] - h["00"])
                           + 14 / 3 * (1 - x))
            - 4 * CF * nf * (pqq(x) * (5 / 9 + h["0"] / 3) + 2 / 3 * (1 - x))
            + 4 * CF ** 2 * (2 * pqq(x) * (h["10"] - 0.75 * h["0"] + h["2"]) - 2 * pqq(-x) * (Z2 + 2 * h["m10"] - h["00"])
                             - (1 - x) * (1 - 1.5 * h["0"]) - h["0"] - (1 + x) * h["00"]))
    if sign > 0:
        return plus
    return plus + 16 * CF * (CF - CA / 2) * (pqq(-x) * (Z2 + 2 * h["m10"] - h["00"]) - 2 * (1 - x) - (1 + x) * h["0"])


def P1_ps(x, nf):
    h = _H(x)
    return 4 * CF * nf * (20 / 9 / x - 2 + 6 * x - 4 * h["0"] + x * x * (8 / 3 * h["0"] - 56 / 9)
                          + (1 + x) * (5 * h["0"] - 2 * h["00"]))


def P1_qg(x, nf):
    h = _H(x)
    return (4 * CA * nf * (20 / 9 / x - 2 + 25 * x - 2 * pqg(-x) * h["m10"] - 2 * pqg(x) * h["11"]
                           + x * x * (44 / 3 * h["0"] - 218 / 9) + 4 * (1 - x) * (h["00"] - 2 * h["0"] + x * h["1"])
                           - 4 * Z2 * x - 6 * h["00"] + 9 * h["0"])
            + 4 * CF * nf * (2 * pqg(x) * (h["10"] + h["11"] + h["2"] - Z2) + 4 * x * x * (h["0"] + h["00"] + 2.5)
                             + 2 * (1 - x) * (h["0"] + h["00"] - 2 * x * h["1"] + 29 / 4) - 7.5 - h["00"] - 0.5 * h["0"]))


def P1_gq(x, nf):
    h = _H(x)
    return (4 * CA * CF * (1 / x + 2 * pgq(x) * (h["10"] + h["11"] + h["2"] - 11 / 6 * h["1"]) - x * x * (8 / 3 * h["0"] - 44 / 9)
                           + 4 * Z2 - 2 - 7 * h["0"] + 2 * h["00"] - 2 * h["1"] * x + (1 + x) * (2 * h["00"] - 5 * h["0"] + 37 / 9)
                           - 2 * pgq(-x) * h["m10"])
            - 4 * CF * nf * (2 / 3 * x - pgq(x) * (2 / 3 * h["1"] - 10 / 9))
            + 4 * CF ** 2 * (pgq(x) * (3 * h["1"] - 2 * h["11"]) + (1 + x) * (h["00"] - 3.5 + 3.5 * h["0"]) - 3 * h["00"] + 1
                             - 1.5 * h["0"] + 2 * h["1"] * x))


def P1_gg(x, nf):
    h = _H(x)
    return (4 * CA * nf * (1 - x - 10 / 9 * pgg(x) - 13 / 9 * (1 / x - x * x) - 2 / 3 * (1 + x) * h["0"])
            + 4 * CA ** 2 * (27 + (1 + x) * (11 / 3 * h["0"] + 8 * h["00"] - 13.5) + 2 * pgg(-x) * (h["00"] - 2 * h["m10"] - Z2)
                             - 67 / 9 * (1 / x - x * x) - 12 * h["0"] - 44 / 3 * x * x * h["0"]
                             + 2 * pgg(x) * (67 / 18 - Z2 + h["00"] + 2 * h["10"] + 2 * h["2"]))
            + 4 * CF * nf * (2 * h["0"] + 2 / 3 / x + 10 / 3 * x * x - 12 + (1 + x) * (4 - 5 * h["0"] - 2 * h["00"])))


# ---- NLO polarised pure singlet, x < 1 ----
def dP1_ps(x, nf):
    lx = np.log(x)
    return 8 * CF * TF * nf * ((1 - x) - (1 - 3 * x) * lx - (1 + x) * lx * lx)


# ---- NLO polarised singlet, x < 1 (Mertig-van Neerven 1996, Vogelsang 1996), normalisation (alpha_s / 2 pi)^2 ----
def _S2(x):
    return -2 * _li2(-x) - 2 * np.log(x) * np.log(1 + x) + 0.5 * np.log(x) ** 2 - np.pi ** 2 / 6


def _dpqg(x):
    return 2 * x - 1


def _dpgq(x):
    return 2 - x


def _dpgg(x):
    return 1 / (1 - x) - 2 * x + 1


def dP1_qg(x, nf):
    lx, l1 = np.log(x), np.log(1 - x)
    return nf * (CF * TF * (-22 + 27 * x - 9 * lx + 8 * (1 - x) * l1 + _dpqg(x) * (2 * l1 ** 2 - 4 * l1 * lx + lx ** 2 - 2 * np.pi ** 2 / 3))
                 + CA * TF * (2 * (12 - 11 * x) - 8 * (1 - x) * l1 + 2 * (1 + 8 * x) * lx - 2 * (l1 ** 2 - np.pi ** 2 / 6) * _dpqg(x)
                              - (2 * _S2(x) - 3 * lx ** 2) * _dpqg(-x)))


def dP1_gq_brackets(x):
    """(CF TF nf, CF^2, CF CA) brackets of Delta P_gq^(1); their first moments are -1/3, -9/4, 71/12
    (Delta gamma_gq^(1)(N=1) = 18 CF^2 - 142/3 CF CA + 8/3 CF TF nf in the gamma = -8 P convention)"""
    lx, l1 = np.log(x), np.log(1 - x)
    b_nf = -4 / 9 * (x + 4) - 4 / 3 * _dpgq(x) * l1
    b_ff = -0.5 - 0.5 * (4 - x) * lx - _dpgq(-x) * l1 + (-4 - l1 ** 2 + 0.5 * lx ** 2) * _dpgq(x)
    b_fa = ((4 - 13 * x) * lx + (10 + x) / 3 * l1 + (41 + 35 * x) / 9 + 0.5 * (-2 * _S2(x) + 3 * lx ** 2) * _dpgq(-x)
            + (l1 ** 2 - 2 * l1 * lx - np.pi ** 2 / 6) * _dpgq(x))
    return b_nf, b_ff, b_fa


def dP1_gq(x, nf, cf2_factor=1.0):
    b_nf, b_ff, b_fa = dP1_gq_brackets(x)
    return CF * TF * nf * b_nf + cf2_factor * CF ** 2 * b_ff + CF * CA * b_fa


def dP1_gg(x, nf):
    lx, l1 = np.log(x), np.log(1 - x)
    return (-CA * TF * nf * (4 * (1 - x) + 4 / 3 * (1 + x) * lx + 20 / 9 * _dpgg(x))
            - CF * TF * nf * (10 * (1 - x) + 2 * (5 - x) * lx + 2 * (1 + x) * lx ** 2)
            + CA ** 2 * ((29 - 67 * x) / 3 * lx - 9.5 * (1 - x) + 4 * (1 + x) * lx ** 2 - 2 * _S2(x) * _dpgg(-x)
                         + (67 / 9 - 4 * l1 * lx + lx ** 2 - np.pi ** 2 / 3) * _dpgg(x)))
