"""`dummy` object replacing the f2py module of tiktaalik/f90wrap/model.f90:16-158 (GK model wrappers)."""
import numpy as np
import torch

from .. import _lib

_WHICH = {"hu": 0, "hd": 1, "hs": 2, "hg": 3, "hu_val": 4, "hd_val": 5, "hu_sea": 6, "hd_sea": 7}


def _dev1(a, dev):
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=torch.float64).reshape(-1).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)).to(dev)


class _Dummy:
    @staticmethod
    def _call(name, x, xi, t):
        dev = _lib.require_cuda()
        x, xi, t = _dev1(x, dev), _dev1(xi, dev), _dev1(t, dev)
        out = torch.empty((x.numel(), xi.numel(), t.numel()), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().tkb_gk(_WHICH[name], x.numel(), xi.numel(), t.numel(), _lib.ptr(x), _lib.ptr(xi),
                                     _lib.ptr(t), _lib.ptr(out), _lib.stream_ptr()))
        return out


for _n in _WHICH:
    setattr(_Dummy, _n + "_wrap", staticmethod(lambda x, xi, t, _n=_n: _Dummy._call(_n, x, xi, t)))

dummy = _Dummy()
