"""`dummy` object replacing the f2py module of tiktaalik/f90wrap/matrices.f90:21-193.

Same method names and argument meaning as f2py generated them (lower-cased names, array
lengths optional).  Arrays that the Fortran wrappers returned as fresh numpy arrays come back
as torch CUDA float64 tensors with the same logical shape: M[i, j, xi, Q2], K[i, j, xi].
By default a getter returns a private copy (f2py intent(out) semantics); with
`zero_copy = True` it returns a view of library memory that is valid until the next
make_kernels_wrap / make_matrices_wrap.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib

_TYPES = {"vns": 0, "vsg": 1, "ans": 2, "asg": 3}
_KTYPES = {"vqq": 0, "vqg": 1, "vgq": 2, "vgg": 3, "aqq": 4, "aqg": 5, "agq": 6, "agg": 7}


def _host_f64(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)


class _Dummy:
    zero_copy = False

    # ---- pixelspace_wrap (f90wrap/matrices.f90:21-30) ----
    @staticmethod
    def pixelspace_wrap(nx, xi, grid_type):
        x = np.empty(int(nx), dtype=np.float64)
        _lib.check(_lib.lib().tkb_pixelspace(int(nx), float(xi), int(grid_type), x.ctypes.data_as(_lib._dp)))
        return x

    # ---- make_kernels_wrap (:35-41) ----
    @staticmethod
    def make_kernels_wrap(nx, xi_array, grid_type, nxi=None):
        _lib.require_cuda()
        xi = _host_f64(xi_array)
        n = xi.shape[0] if nxi is None else int(nxi)
        _lib.check(_lib.lib().tkb_make_kernels(int(nx), n, xi.ctypes.data_as(_lib._dp), int(grid_type),
                                               _lib.stream_ptr()))

    @staticmethod
    def make_kernels_shard(nx, xi_array, grid_type, clamp_first, clamp_last):
        _lib.require_cuda()
        xi = _host_f64(xi_array)
        _lib.check(_lib.lib().tkb_make_kernels_shard(int(nx), xi.shape[0], xi.ctypes.data_as(_lib._dp),
                                                     int(grid_type), int(clamp_first), int(clamp_last),
                                                     _lib.stream_ptr()))

    @staticmethod
    def set_variant(eta_recip_f32=False, x0_recip_f32=False):
        """-Ofast float32-division variant of the node formula / stencil origin (not an f2py name)."""
        _lib.require_cuda()
        _lib.check(_lib.lib().tkb_set_variant(int(bool(eta_recip_f32)), int(bool(x0_recip_f32))))

    # ---- make_matrices_wrap (:43-53) ----
    @staticmethod
    def make_matrices_wrap(q2_array, l_nlo, nq2=None):
        _lib.require_cuda()
        q2 = _host_f64(q2_array)
        n = q2.shape[0] if nq2 is None else int(nq2)
        _lib.check(_lib.lib().tkb_make_matrices(n, q2.ctypes.data_as(_lib._dp), int(bool(l_nlo)), _lib.stream_ptr()))

    # ---- get_nx_wrap / get_nxi_wrap / get_nQ2_wrap (:58-71) ----
    @staticmethod
    def get_nx_wrap():
        return _lib.lib().tkb_get_nx()

    @staticmethod
    def get_nxi_wrap():
        return _lib.lib().tkb_get_nxi()

    @staticmethod
    def get_nq2_wrap():
        return _lib.lib().tkb_get_nq2()

    # ---- evomatrix_*_wrap (:76-110) ----
    def _evomatrix(self, kind, nx, nxi, nq2, nstype):
        L = _lib.lib()
        dev = _lib.require_cuda()
        if (nx, nxi, nq2) != (L.tkb_get_nx(), L.tkb_get_nxi(), L.tkb_get_nq2()):
            raise RuntimeError("evomatrix_%s_wrap: (nx, nxi, nQ2) do not match the cached matrices" % kind)
        m = nx * (2 if kind in ("vsg", "asg") else 1)
        if self.zero_copy:
            addr = C.c_void_p()
            shape = (C.c_int64 * 4)()
            strides = (C.c_int64 * 4)()
            _lib.check(L.tkb_evomatrix(_TYPES[kind], int(nstype), C.byref(addr), shape, strides))
            store = _lib.view_tensor(addr.value, (nxi, nq2, m, m))
        else:
            store = torch.empty((nxi, nq2, m, m), dtype=torch.float64, device=dev)
            _lib.check(L.tkb_evomatrix_copy(_TYPES[kind], int(nstype), _lib.ptr(store), _lib.stream_ptr()))
        return store.permute(2, 3, 0, 1)          # logical (i, j, xi, Q2)

    def evomatrix_vns_wrap(self, nx, nxi, nq2, nstype):
        return self._evomatrix("vns", nx, nxi, nq2, nstype)

    def evomatrix_ans_wrap(self, nx, nxi, nq2, nstype):
        return self._evomatrix("ans", nx, nxi, nq2, nstype)

    def evomatrix_vsg_wrap(self, nx, nxi, nq2):
        return self._evomatrix("vsg", nx, nxi, nq2, 1)

    def evomatrix_asg_wrap(self, nx, nxi, nq2):
        return self._evomatrix("asg", nx, nxi, nq2, 1)

    # ---- evokernel_*_wrap (:115-193) ----
    @staticmethod
    def _evokernel(kind, q2, nx, nxi, nfl, l_nlo, i_ns_type=1):
        L = _lib.lib()
        dev = _lib.require_cuda()
        if (nx, nxi) != (L.tkb_get_nx(), L.tkb_get_nxi()):
            raise RuntimeError("evokernel_%s_wrap: (nx, nxi) do not match the cached kernels" % kind)
        store = torch.empty((nxi, nx, nx), dtype=torch.float64, device=dev)
        _lib.check(L.tkb_evokernel(_KTYPES[kind], float(q2), int(nfl), int(bool(l_nlo)), int(i_ns_type),
                                   _lib.ptr(store), _lib.stream_ptr()))
        return store.permute(1, 2, 0)             # logical (i, j, xi)

    def evokernel_vqq_wrap(self, q2, nx, nxi, nfl, l_nlo, i_ns_type):
        return self._evokernel("vqq", q2, nx, nxi, nfl, l_nlo, i_ns_type)

    def evokernel_vqg_wrap(self, q2, nx, nxi, nfl, l_nlo):
        return self._evokernel("vqg", q2, nx, nxi, nfl, l_nlo)

    def evokernel_vgq_wrap(self, q2, nx, nxi, nfl, l_nlo):
        return self._evokernel("vgq", q2, nx, nxi, nfl, l_nlo)

    def evokernel_vgg_wrap(self, q2, nx, nxi, nfl, l_nlo):
        return self._evokernel("vgg", q2, nx, nxi, nfl, l_nlo)

    def evokernel_aqq_wrap(self, q2, nx, nxi, nfl, l_nlo, i_ns_type):
        return self._evokernel("aqq", q2, nx, nxi, nfl, l_nlo, i_ns_type)

    def evokernel_aqg_wrap(self, q2, nx, nxi, nfl, l_nlo):
        return self._evokernel("aqg", q2, nx, nxi, nfl, l_nlo)

    def evokernel_agq_wrap(self, q2, nx, nxi, nfl, l_nlo):
        return self._evokernel("agq", q2, nx, nxi, nfl, l_nlo)

    def evokernel_agg_wrap(self, q2, nx, nxi, nfl, l_nlo):
        return self._evokernel("agg", q2, nx, nxi, nfl, l_nlo)

    # ---- application: the np.einsum the reference's callers run on the matrices ----
    @staticmethod
    def apply_wrap(M, g, out=None):
        """out[i, k, ..., q] = sum_j M[i, j, k, q] g[j, k, ...]   (evolution.py:259, examples.py:77).

        M: (m, m, nxi, nQ2) as returned by the evomatrix getters (any strides);
        g: (m, nxi, ...) numpy array or torch tensor; a block of samples cut out of a larger (m, nxi, S_total)
        device array (g_all[:, :, s0:s1]) is read in place.  out: optional contiguous CUDA storage of shape
        (nxi, nQ2, m, S) to write into (an inference loop reuses it).  Returns a CUDA tensor (m, nxi, ..., nQ2).
        """
        L = _lib.lib()
        dev = _lib.require_cuda()
        if not isinstance(M, torch.Tensor):
            M = torch.as_tensor(np.asarray(M), dtype=torch.float64)
        M = M.to(device=dev, dtype=torch.float64)
        if not isinstance(g, torch.Tensor):
            g = torch.as_tensor(np.ascontiguousarray(g), dtype=torch.float64)
        g = g.to(device=dev, dtype=torch.float64)
        m, m2, nxi, nq2 = M.shape
        if m != m2 or g.shape[0] != m or g.shape[1] != nxi:
            raise ValueError("apply: shapes %s and %s do not match 'ijkl,jk...->ik...l'" % (tuple(M.shape), tuple(g.shape)))
        extra = tuple(g.shape[2:])
        S = 1
        for e in extra:
            S *= int(e)
        store = M.permute(2, 3, 0, 1)             # [xi][Q2][i][j]
        if not store.is_contiguous():
            store = store.contiguous()
        if g.dim() == 3 and S > 0 and (g.stride(2) == 1 or S == 1) and g.stride(0) >= 0 and g.stride(1) >= 0:
            g2 = g                                # sample block of a larger array: strides go to the library
        else:
            g2 = g.reshape(m, nxi, S)
            if not g2.is_contiguous():
                g2 = g2.contiguous()
        if out is None:
            out = torch.empty((nxi, nq2, m, S), dtype=torch.float64, device=dev)
        elif (not isinstance(out, torch.Tensor) or tuple(out.shape) != (nxi, nq2, m, S) or out.dtype != torch.float64
              or not out.is_cuda or not out.is_contiguous()):
            raise ValueError("apply: out must be a contiguous float64 CUDA tensor of shape %s" % ((nxi, nq2, m, S),))
        if S > 0:
            _lib.check(L.tkb_apply_strided(int(m), int(nxi), int(nq2), int(S), _lib.ptr(store), _lib.ptr(g2),
                                           int(g2.stride(0)), int(g2.stride(1)), _lib.ptr(out), _lib.stream_ptr()))
        return out.permute(2, 0, 3, 1).reshape((m, nxi) + extra + (nq2,))


dummy = _Dummy()
