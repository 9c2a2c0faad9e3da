"""
evolution.py -- the Evolver convenience class and its helpers, same API as the reference's
tiktaalik/evolution.py:19-548, with every array on the B200.

The matrix applications (np.einsum in the reference, evolution.py:259-348, 507-509) run through
the library's apply kernel (tkb_apply); the flavour-basis rotations are torch device ops.
GPD grids may be numpy arrays or torch tensors; results are torch CUDA float64 tensors with
the reference's shapes, e.g. evolveGPDs -> (6, nx, nxi, ..., nQ2).

Reference behaviours kept on purpose, switchable with
`evolution.REFERENCE_BUGS = False`:
  * Evolver.x is always the *linear* pixelspace, also for grid_type=2 (evolution.py:104);
  * evolveSinglet's final-segment selector tests `gpd_type=='V' and nlo==False` twice, so
    V/NLO never selects Mevo_VSG_NLO_Q2f (evolution.py:273-280): it reuses the previous
    segment's matrix, or fails if no threshold was passed.
"""

import numpy as np
import torch

from . import matrices
from . import pars
from . import _lib
from .matrices import pixelspace, Q2space, passes, get_nfl, apply_matrix

REFERENCE_BUGS = True

__all__ = ["Evolver", "make_evo_basis", "make_flv_basis", "evolve_fixed_flavor", "Tpls_matrix",
           "add_pad", "swap_AV"]


def _dev(a):
    """numpy / torch -> CUDA float64 tensor"""
    dev = _lib.require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=torch.float64)
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(dev)


def _apply_flavoured(M, g):
    """np.einsum('ijkq,fjk...->fik...q', M, g): the flavour axis rides along as extra samples."""
    nd = g.dim()
    moved = g.permute(*range(1, 3), 0, *range(3, nd))        # (j, k, f, ...)
    out = apply_matrix(M, moved)                              # (i, k, f, ..., q)
    nd_o = out.dim()
    return out.permute(2, 0, 1, *range(3, nd_o))              # (f, i, k, ..., q)


# ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
# Create a class to cache matrices and properties into

_SEGMENTS = ("mc2", "mb2", "Q2f")
_KINDS = ("VNS", "ANS", "VSG", "ASG")


class Evolver:
    ''' A convenience class for evolving GPDs (reference: evolution.py:19-117).
        Darwin = Evolver(nx=100, xi=0.5, nQ2=17, Q2i=mc2+eps, Q2f=mb2-eps, grid_type=1)
    - nx .......... even number of x points (>= 6; divisible by 4 for grid_type=2)
    - xi .......... float or array of xi values
    - nQ2 ......... number of Q2 points (use >= 4 so both thresholds can be injected)
    - Q2i, Q2f .... initial and final Q2
    - grid_type ... 1 (linear) or 2 (log-linear-log)
    The Q2 grid is matrices.Q2space(Q2i, Q2f, nQ2).  Up to 24 matrices are cached as attributes
    Mevo_{VNS,ANS,VSG,ASG}_{LO,NLO}_{mc2,mb2,Q2f}; at NLO "VNS" holds the plus-type and "ANS" the
    minus-type non-singlet matrix.
    '''
    pass_charm  = False
    pass_bottom = False

    def __init__(self,
            nx=100,
            xi=0.5,
            nQ2=17,
            Q2i=pars.mc2+pars.epsilon,
            Q2f=pars.mb2-pars.epsilon,
            grid_type=1
            ):
        for seg in _SEGMENTS:
            for order in ("LO", "NLO"):
                for kind in _KINDS:
                    setattr(self, "Mevo_%s_%s_%s" % (kind, order, seg), None)
        try:
            assert(2*(nx//2)==nx)
            assert(nx>=6)
        except:
            raise ValueError("nx must be even, not nx={:d}".format(nx))
        if(np.isscalar(xi)):
            xi = np.array([xi])
        xi = np.asarray(xi, dtype=np.float64)
        self.nx  = nx
        self.nxi = xi.shape[0]
        self.grid_type = grid_type
        # reference: always the linear grid here (evolution.py:104)
        self.x   = pixelspace(self.nx) if REFERENCE_BUGS or self.nxi != 1 else pixelspace(self.nx, xi[0], grid_type)
        self.xi  = xi
        matrices.initialize_kernels(nx, xi, grid_type=grid_type)
        self.Q2i = Q2i
        self.Q2f = Q2f
        self.nQ2 = nQ2
        self.Q2 = Q2space(Q2i, Q2f, nQ2)
        # We use (up to) five flavors internally
        self.nfl = 5
        self.makeEvolutionMatrices()
        return

    # -- matrix selection ----------------------------------------------------------------

    def _triplet(self, seg, gpd_type, nlo):
        """(M_SG, M_NS+, M_NS-) for one threshold segment (reference: evolution.py:156-171)."""
        order = "NLO" if nlo else "LO"
        get = lambda kind: getattr(self, "Mevo_%s_%s_%s" % (kind, order, seg))
        if gpd_type == 'V':
            return get("VSG"), get("VNS"), (get("ANS") if nlo else get("VNS"))
        return get("ASG"), get("ANS"), (get("VNS") if nlo else get("ANS"))

    ## ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~

    def evolveGPDs(self, gpd_grid, gpd_type='V', nlo=False):
        ''' Evolve a GPD grid given in the physical flavour basis (reference: evolution.py:121-213).
        - gpd_grid ... dim(nfl+1, nx, nxi, ...), flavours ordered (glue, up, down, ...),
                       x values on pixelspace(nx)
        - gpd_type ... 'V' or 'A';  nlo ... False or True
        Returns dim(6, nx, nxi, ..., nQ2); missing heavy flavours are zero.
        '''
        gpd_grid = _dev(gpd_grid)
        if(gpd_grid.dim() < 3):
            gpd_grid = gpd_grid[...,None]
        assert(self.nx  == gpd_grid.shape[1])
        assert(self.nxi == gpd_grid.shape[2])
        assert( (gpd_type=='V') or (gpd_type=='A') )
        # Segments: (below mc2, nfl=3), (mc2..mb2, nfl=4), (.. Q2f, nfl at Q2f); each starts from the last Q2 point of
        # the previous one and contributes all its Q2 points but the first (reference: evolution.py:121-150).
        segs = []
        if(self.pass_charm):
            segs.append(("mc2", 3))
        if(self.pass_bottom):
            segs.append(("mb2", 4))
        segs.append(("Q2f", get_nfl(self.Q2f)))
        trip = [self._triplet(seg, gpd_type, nlo) for seg, _ in segs]
        F, nx, nxi = (int(v) for v in gpd_grid.shape[:3])
        extra = tuple(int(v) for v in gpd_grid.shape[3:])
        S = 1
        for e in extra:
            S *= e
        Fout = max(F, self.nfl + 1)
        nq_tot = 1 + sum(int(t[0].shape[3]) - 1 for t in trip)
        # stored [flavour][x][xi][Q2][sample] (sample fastest: coalesced in the kernels), returned as a permuted view
        out = torch.empty((Fout, nx, nxi, nq_tot, S), dtype=torch.float64, device=gpd_grid.device)
        cur = torch.zeros((Fout, nx, nxi, S), dtype=torch.float64, device=gpd_grid.device)
        cur[:F] = gpd_grid.reshape(F, nx, nxi, S)
        out[:, :, :, 0, :] = cur
        q0 = 1
        for (seg, nfl), (M_SG, M_NSp, M_NSm) in zip(segs, trip):
            nq = int(M_SG.shape[3])
            if S > 0 and nq > 1:
                _fused_fixed_flavor(cur, M_SG, M_NSp, M_NSm, nfl, out, q0, 1)
            q0 += nq - 1
            cur = out[:, :, :, q0 - 1, :].contiguous()
        new_gpd = out.permute(0, 1, 2, 4, 3).reshape((Fout, nx, nxi) + extra + (nq_tot,))
        return new_gpd

    ## ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~

    def evolveSinglet(self, singlet_grid, gluon_grid, gpd_type='V', nlo=False):
        ''' Evolve only a singlet-quark grid and a gluon grid, each dim(nx, nxi, ...)
        (reference: evolution.py:217-287).  Returns (singlet, gluon), each dim(nx, nxi, ..., nQ2). '''
        singlet_grid = _dev(singlet_grid)
        gluon_grid = _dev(gluon_grid)
        if(singlet_grid.dim() < 2):
            singlet_grid = singlet_grid[...,None]
        if(gluon_grid.dim() < 2):
            gluon_grid = gluon_grid[...,None]
        assert((self.nx ==singlet_grid.shape[0]) and (self.nx ==gluon_grid.shape[0]))
        assert((self.nxi==singlet_grid.shape[1]) and (self.nxi==gluon_grid.shape[1]))
        assert( (gpd_type=='V') or (gpd_type=='A') )
        sg_grid = torch.cat((singlet_grid, gluon_grid), dim=0)
        new_gpd = sg_grid.clone()[...,None]
        M_SG = None
        if(self.pass_charm):
            M_SG = self._triplet("mc2", gpd_type, nlo)[0]
            added_gpd = apply_matrix(M_SG, new_gpd[...,-1])
            new_gpd = torch.cat((new_gpd, added_gpd[...,1:]), dim=-1)
        if(self.pass_bottom):
            M_SG = self._triplet("mb2", gpd_type, nlo)[0]
            added_gpd = apply_matrix(M_SG, new_gpd[...,-1])
            new_gpd = torch.cat((new_gpd, added_gpd[...,1:]), dim=-1)
        if REFERENCE_BUGS and gpd_type == 'V' and nlo:
            # evolution.py:273-280: no branch matches, M_SG keeps its previous value
            if M_SG is None:
                raise UnboundLocalError("cannot access local variable 'M_SG' where it is not associated with a value")
        else:
            M_SG = self._triplet("Q2f", gpd_type, nlo)[0]
        added_gpd = apply_matrix(M_SG, new_gpd[...,-1])
        new_gpd = torch.cat((new_gpd, added_gpd[...,1:]), dim=-1)
        new_singlet = new_gpd[0:self.nx,...]
        new_gluon = new_gpd[self.nx:2*self.nx,...]
        return new_singlet, new_gluon

    def evolveNS(self, ns_grid, gpd_type='V', nlo=False, ns_type=1):
        ''' Evolve only a non-singlet quark grid dim(nx, nxi, ...) (reference: evolution.py:289-350).
        ns_type = 1 (plus-type) or -1 (minus-type).  Returns dim(nx, nxi, ..., nQ2). '''
        ns_grid = _dev(ns_grid)
        if(ns_grid.dim() < 2):
            ns_grid = ns_grid[...,None]
        assert(self.nx ==ns_grid.shape[0])
        assert(self.nxi==ns_grid.shape[1])
        assert( (gpd_type=='V') or (gpd_type=='A') )
        assert( (ns_type==-1) or (ns_type==1) )
        # V+ = A-, V- = A+
        if(ns_type==-1):
            gpd_type = swap_AV(gpd_type)
        order = "NLO" if nlo else "LO"
        kind = "VNS" if gpd_type == 'V' else "ANS"
        new_gpd = ns_grid.clone()[...,None]
        segs = (["mc2"] if self.pass_charm else []) + (["mb2"] if self.pass_bottom else []) + ["Q2f"]
        for seg in segs:
            M_NS = getattr(self, "Mevo_%s_%s_%s" % (kind, order, seg))
            added_gpd = apply_matrix(M_NS, new_gpd[...,-1])
            new_gpd = torch.cat((new_gpd, added_gpd[...,1:]), dim=-1)
        return new_gpd

    ## ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~

    def makeEvolutionMatrices(self):
        ''' Build and cache the evolution matrices, one set per threshold segment
        (reference: evolution.py:354-379). '''
        self.pass_charm = False
        self.pass_bottom = False
        if(passes(self.Q2,pars.mc2)):
            self.pass_charm = True
            subQ2 = self.Q2[self.Q2 <= pars.mc2]
            self.makeEvolutionMatrices_mc2(subQ2)
        if(passes(self.Q2,pars.mb2)):
            self.pass_bottom = True
            subQ2 = self.Q2[(pars.mc2 <= self.Q2) & (self.Q2 <= pars.mb2)]
            self.makeEvolutionMatrices_mb2(subQ2)
        if(self.pass_bottom):
            subQ2 = self.Q2[pars.mb2 <= self.Q2]
            self.makeEvolutionMatrices_Q2f(subQ2)
        elif(self.pass_charm):
            subQ2 = self.Q2[pars.mc2 <= self.Q2]
            self.makeEvolutionMatrices_Q2f(subQ2)
        else:
            self.makeEvolutionMatrices_Q2f(self.Q2)
        self.initialized = True
        return

    def _make_segment(self, seg, Q2):
        # LO first, then NLO (reference: evolution.py:381-429).  The cached matrices outlive the next
        # initialize_evolution_matrices call, which rewrites the library's storage in place, so they are always
        # private copies -- also when the getters are switched to zero-copy views (f90wrap.matrices.dummy.zero_copy).
        f90 = matrices.f90src
        keep = getattr(f90, "zero_copy", False)
        f90.zero_copy = False
        try:
            for order, nlo in (("LO", False), ("NLO", True)):
                matrices.initialize_evolution_matrices(Q2, nlo=nlo)
                setattr(self, "Mevo_VNS_%s_%s" % (order, seg), matrices.matrix_VNS())
                setattr(self, "Mevo_ANS_%s_%s" % (order, seg), matrices.matrix_ANS())
                setattr(self, "Mevo_VSG_%s_%s" % (order, seg), matrices.matrix_VSG())
                setattr(self, "Mevo_ASG_%s_%s" % (order, seg), matrices.matrix_ASG())
        finally:
            f90.zero_copy = keep

    def makeEvolutionMatrices_mc2(self,Q2):
        ''' Matrices up to mc2, if the charm threshold is passed. '''
        self._make_segment("mc2", Q2)

    def makeEvolutionMatrices_mb2(self,Q2):
        ''' Matrices from mc2 to mb2, if the bottom threshold is passed. '''
        self._make_segment("mb2", Q2)

    def makeEvolutionMatrices_Q2f(self,Q2):
        ''' Matrices up to the target Q2. '''
        self._make_segment("Q2f", Q2)

# ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
# Helpful routines used by the class


def make_evo_basis(gpd_grid, nfl):
    ''' At fixed nfl, convert a flavour-basis grid dim(nfl+1, nx, nxi, ...) ordered
    (glue, up, down, ...) into (reference: evolution.py:435-462)
        1. sg_grid  .... dim(2*nx, nxi, ...), singlet then gluon
        2. t_grid ...... dim(nfl-1, nx, nxi, ...), (T3, T8, ...)
        3. qmin_grid ... dim(nfl, nx, nxi, ...), (umin, dmin, ...)
    '''
    gpd_work = _dev(gpd_grid).clone()
    if(gpd_work.shape[0] < nfl+1):
        gpd_work = add_pad(gpd_work, nfl+1 - gpd_work.shape[0])
    q = gpd_work[1:nfl+1,...]
    qflip = torch.flip(q, dims=(1,))
    qmin_grid = q + qflip
    qpls_grid = q - qflip
    M = torch.as_tensor(Tpls_matrix(nfl), dtype=torch.float64, device=gpd_work.device)
    st_grid = torch.einsum('ij,j...->i...', M, qpls_grid)
    t_grid = st_grid[1:,...]
    s_grid = st_grid[0,...]
    g_grid = gpd_work[0,...]
    sg_grid = torch.cat((s_grid, g_grid), dim=0)
    return sg_grid, t_grid, qmin_grid


def make_flv_basis(sg_grid, t_grid, qmin_grid, nfl):
    ''' Inverse of make_evo_basis: back to dim(nfl+1, nx, nxi, ...) ordered (glue, up, down, ...)
    (reference: evolution.py:465-492). '''
    nx = t_grid.shape[1]
    M = Tpls_matrix(nfl)
    N = torch.as_tensor(np.linalg.inv(M), dtype=torch.float64, device=sg_grid.device)
    s_grid = sg_grid[None,0:nx,...]
    st_grid = torch.cat((s_grid, t_grid), dim=0)
    qpls_grid = torch.einsum('ij,j...->i...', N, st_grid)
    q_grid = (qpls_grid + qmin_grid) / 2
    g_grid = sg_grid[None,nx:2*nx,...]
    gpd_grid = torch.cat((g_grid, q_grid), dim=0)
    return gpd_grid


def _storage(M):
    """[xi][Q2][i][j] storage of a matrix of logical shape (m, m, nxi, nQ2); no copy for what the getters return"""
    st = M.permute(2, 3, 0, 1)
    return st if st.is_contiguous() else st.contiguous()


def _fused_fixed_flavor(gpd_in, M_SG, M_NSp, M_NSm, nfl, out, q0, skip):
    """tkb_evolve_fixed_flavor: gpd_in (F, nx, nxi, S) contiguous -> out[:, :, :, q0 : q0 + nQ2 - skip, :]
    (out is stored [flavour][x][xi][Q2][sample])"""
    import ctypes as C
    L = _lib.lib()
    F, nx, nxi, S = (int(v) for v in gpd_in.shape)
    nq = int(M_SG.shape[3])
    T = np.ascontiguousarray(Tpls_matrix(nfl), dtype=np.float64)
    N = np.ascontiguousarray(np.linalg.inv(T), dtype=np.float64)
    sg, nsp, nsm = _storage(M_SG), _storage(M_NSp), _storage(M_NSm)
    nwork = int(L.tkb_evolve_workspace(nx, nxi, nq, S, nfl))
    work = torch.empty((max(nwork, 1),), dtype=torch.float64, device=gpd_in.device)
    _lib.check(L.tkb_evolve_fixed_flavor(nx, nxi, nq, S, nfl, F, int(out.shape[0]), T.ctypes.data_as(_lib._dp),
                                         N.ctypes.data_as(_lib._dp), _lib.ptr(sg), _lib.ptr(nsp), _lib.ptr(nsm),
                                         _lib.ptr(gpd_in), _lib.ptr(out), int(out.shape[3]), int(q0), int(skip),
                                         _lib.ptr(work), _lib.stream_ptr()))


def evolve_fixed_flavor(gpd_grid, M_SG, M_NSp, M_NSm, nfl=4):
    ''' Evolve a flavour-basis grid with given singlet / non-singlet matrices at a fixed number of
    flavours: decouple, apply, recouple (reference: evolution.py:495-514).  One fused library call
    (csrc/evolve.cu); evolve_fixed_flavor_unfused is the same pipeline as separate tensor operations. '''
    gpd_grid = _dev(gpd_grid)
    M_SG, M_NSp, M_NSm = _dev(M_SG), _dev(M_NSp), _dev(M_NSm)
    F, nx, nxi = (int(v) for v in gpd_grid.shape[:3])
    extra = tuple(int(v) for v in gpd_grid.shape[3:])
    S = 1
    for e in extra:
        S *= e
    nq = int(M_SG.shape[3])
    Fout = max(F, nfl + 1)
    out = torch.empty((Fout, nx, nxi, nq, S), dtype=torch.float64, device=gpd_grid.device)
    if S > 0:
        _fused_fixed_flavor(gpd_grid.reshape(F, nx, nxi, S).contiguous(), M_SG, M_NSp, M_NSm, nfl, out, 0, 0)
    # stored with the sample index fastest; the logical shape is the reference's (flavour, x, xi, ..., Q2)
    return out.permute(0, 1, 2, 4, 3).reshape((Fout, nx, nxi) + extra + (nq,))


def evolve_fixed_flavor_unfused(gpd_grid, M_SG, M_NSp, M_NSm, nfl=4):
    ''' evolve_fixed_flavor as the reference spells it: make_evo_basis, three applications, make_flv_basis,
    padding (reference: evolution.py:495-514). '''
    gpd_grid = _dev(gpd_grid)
    extra_nfl = gpd_grid.shape[0] - 1 - nfl
    sg_ini, t_ini, qmin_ini = make_evo_basis(gpd_grid, nfl=nfl)
    sg_fin   = apply_matrix(M_SG, sg_ini)
    t_fin    = _apply_flavoured(M_NSp, t_ini)
    qmin_fin = _apply_flavoured(M_NSm, qmin_ini)
    gpd_evo = make_flv_basis(sg_fin, t_fin, qmin_fin, nfl=nfl)
    gpd_evo = add_pad(gpd_evo, extra_nfl)
    return gpd_evo


def Tpls_matrix(nfl):
    ''' (u+, d+, ...) -> (Sigma, T3, T8, ...); invert to go back (reference: evolution.py:517-529).
    Small host matrix. '''
    M = np.zeros((nfl,nfl))
    M[0,:] = 1
    for ifl in range(1,nfl):
        M[ifl,0:ifl] = 1
        M[ifl,ifl] = -ifl
    return M


def add_pad(grid, nrows):
    ''' Append nrows rows of zeros to dimension 0 (reference: evolution.py:532-540). '''
    if nrows <= 0:
        return grid
    shape = (nrows,) + tuple(grid.shape[1:])
    if isinstance(grid, torch.Tensor):
        return torch.cat((grid, torch.zeros(shape, dtype=grid.dtype, device=grid.device)), dim=0)
    return np.concatenate((grid, np.zeros(shape)), axis=0)


def swap_AV(gpd_type):
    if(gpd_type=='A'):
        return 'V'
    if(gpd_type=='V'):
        return 'A'
    raise ValueError("gpd_type must be 'V' or 'A', not '{}'.".format(gpd_type))
