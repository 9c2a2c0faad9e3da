"""Running coupling (reference: tiktaalik/qcd.py:4-7 -> f90wrap/qcd.f90:19-29)."""
import numpy as np
from .f90wrap.qcd import dummy as f90src

def alphaQCD(Q2):
    if(np.isscalar(Q2)):
        Q2 = np.array([Q2])
    return f90src.alpha_wrap(Q2)
