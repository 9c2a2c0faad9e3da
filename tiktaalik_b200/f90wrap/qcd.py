"""`dummy` object replacing the f2py module of tiktaalik/f90wrap/qcd.f90:19-29."""
import ctypes as C

import numpy as np

from .. import _lib


class _Dummy:
    @staticmethod
    def alpha_wrap(q2):
        q2 = np.ascontiguousarray(np.atleast_1d(q2), dtype=np.float64)
        out = np.empty_like(q2)
        _lib.check(_lib.lib().tkb_alpha(q2.shape[0], q2.ctypes.data_as(_lib._dp), out.ctypes.data_as(_lib._dp)))
        return out


dummy = _Dummy()
