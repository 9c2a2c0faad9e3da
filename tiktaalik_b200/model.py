"""
model.py -- the Goloskokov-Kroll model GPD, same API as the reference's tiktaalik/model.py:22-114
(H_singlet, Hu, Hd, Hs, Hg, Hu_val, Hd_val, Hu_sea, Hd_sea; arguments x, xi, t scalars or 1-d arrays;
result of shape (nx, nxi, nt)), evaluated on the device: results are torch CUDA float64 tensors.
`model.host` is a numpy mirror for machines without a GPU.  See f90src/model/gk.f90 for the model.
"""

import numpy as np
import torch
from .f90wrap.model import dummy as f90src
from . import model_host as host

# ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
# Model GPDs

def _neg(x):
    return -x if isinstance(x, torch.Tensor) else -np.atleast_1d(np.asarray(x, dtype=np.float64))

def H_singlet(x, xi, t):
    f = (
              Hu(x,xi,t) - Hu(_neg(x),xi,t)
            + Hd(x,xi,t) - Hd(_neg(x),xi,t)
            + Hs(x,xi,t) - Hs(_neg(x),xi,t)
            )
    return f

def Hu(x, xi, t):
    return f90src.hu_wrap(x,xi,t)

def Hd(x, xi, t):
    return f90src.hd_wrap(x,xi,t)

def Hs(x, xi, t):
    return f90src.hs_wrap(x,xi,t)

def Hg(x, xi, t):
    return f90src.hg_wrap(x,xi,t)

def Hu_val(x, xi, t):
    return f90src.hu_val_wrap(x,xi,t)

def Hd_val(x, xi, t):
    return f90src.hd_val_wrap(x,xi,t)

def Hu_sea(x, xi, t):
    return f90src.hu_sea_wrap(x,xi,t)

def Hd_sea(x, xi, t):
    return f90src.hd_sea_wrap(x,xi,t)
