"""Constants of the Python layer (reference: tiktaalik/pars.py:2-10).

mc2 and mb2 must stay bit-equal to the compile-time thresholds of the library
(csrc/alpha.cpp; reference constants.f90:30-31): get_neff compares with >= and the
alpha_s shortcuts compare with == (alpha_qcd.f90:29-36, 120-126).
"""
# Masses
mc = 1.29
mb = 4.18
mt = 172.44
mc2 = mc**2
mb2 = mb**2
mt2 = mt**2

# Boop
epsilon = 1e-6
