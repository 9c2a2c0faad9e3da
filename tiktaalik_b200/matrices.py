"""
matrices.py -- kernel and evolution-matrix constructors, same API as the reference's
tiktaalik/matrices.py:18-276 (names, kwargs, asserts, call-order contract), on B200.

Differences that a user can see: arrays are torch CUDA float64 tensors (same logical
shapes: M[i, j, xi, Q2], K[i, j, xi]); numpy inputs are accepted everywhere.
"""

import numpy as np
import torch
from .f90wrap.matrices import dummy as f90src

from . import pars

# ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
# All-important initialization routines.

def initialize_kernels(nx, xi, grid_type=1):
    ''' Initializes all the kernels to be used to make evolution matrices
    (reference: matrices.py:18-46).  This **MUST** be called before any other method.
    - nx .......... even integer >= 6, number of x points
    - xi .......... array of xi points (a scalar becomes a one-component array)
    - grid_type ... 1 linear grid, 2 log / linear / log grid
    Calling it again replaces the kernels and frees any evolution matrices.
    '''
    assert(nx==(nx//2)*2)
    assert(nx>=6)
    if(np.isscalar(xi)):
        xi = np.array([xi])
    if isinstance(xi, torch.Tensor):
        xi = xi.detach().cpu().numpy()
    xi = np.asarray(xi, dtype=np.float64)
    nxi = xi.shape[0]
    f90src.make_kernels_wrap(nx, xi, grid_type)
    return

def initialize_evolution_matrices(Q2, nlo=False):
    ''' Initializes the evolution matrices (reference: matrices.py:48-70).
    - Q2 .... array with *at least* two values; the first is where evolution starts
    - nlo ... whether to include NLO corrections
    initialize_kernels **MUST** have been called first.
    '''
    if isinstance(Q2, torch.Tensor):
        Q2 = Q2.detach().cpu().numpy()
    Q2 = np.asarray(Q2, dtype=np.float64)
    nQ2 = Q2.shape[0]
    assert(nQ2 >= 2)
    f90src.make_matrices_wrap(Q2, nlo)
    return

# ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
# Spaces used

def pixelspace(nx, xi=0.5, grid_type=1):
    ''' The x space used by tiktaalik (reference: matrices.py:75-93): midpoints of nx
    intervals of the eta grid pushed forward by the grid map.  Host numpy array. '''
    x = f90src.pixelspace_wrap(nx, xi, grid_type)
    return x

def Q2space(Q2i, Q2f, nQ2):
    ''' Mostly a geomspace, with the quark mass thresholds injected if they are inside
    (reference: matrices.py:95-106). '''
    nThresh = int(Q2i < pars.mc2 and Q2f > pars.mc2) + int(Q2i < pars.mb2 and Q2f > pars.mb2)
    Q2 = np.geomspace(Q2i, Q2f, nQ2-nThresh)
    if(Q2i < pars.mc2 and Q2f > pars.mc2):
        Q2 = np.append(Q2, pars.mc2)
    if(Q2i < pars.mb2 and Q2f > pars.mb2):
        Q2 = np.append(Q2, pars.mb2)
    Q2 = np.sort(Q2)
    return Q2

# ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
# Q2 related methods

def passes(Q2_array, m2):
    ''' Is m2 strictly inside the (monotonic) Q2_array?  (reference: matrices.py:111-118) '''
    if(Q2_array[0] < m2-pars.epsilon and Q2_array[-1] > m2+pars.epsilon):
        return True
    return False

def get_nfl(Q2):
    ''' Effective number of flavours, 3, 4 or 5 (reference: matrices.py:120-130). '''
    if(Q2 < pars.mc2):
        return 3
    elif(Q2 < pars.mb2):
        return 4
    return 5

# ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
# Routines to create evolution matrices

def _scrub(M):
    # reference: M[np.isnan(M)] = 0; M[np.isinf(M)] = 0 (matrices.py:146-147).  The getters' default result is a
    # private copy and is scrubbed in place like there; a zero-copy view (f90src.zero_copy) is a window on the
    # library's own storage, which must not be rewritten: it is returned as it is when every entry is finite
    # and copied otherwise.
    if getattr(f90src, "zero_copy", False):
        if bool(torch.isfinite(M).all()):
            return M
        return torch.nan_to_num(M, nan=0.0, posinf=0.0, neginf=0.0)
    return torch.nan_to_num_(M, nan=0.0, posinf=0.0, neginf=0.0)

def matrix_VNS(ns_type=1):
    ''' Evolution matrix for *non-singlet* Q->Q, helicity-independent (V-type);
    ns_type = 1 (plus-type) or -1 (minus-type), only matters at NLO (reference: matrices.py:135-148). '''
    nx  = f90src.get_nx_wrap()
    nxi = f90src.get_nxi_wrap()
    nQ2 = f90src.get_nq2_wrap()
    M = f90src.evomatrix_vns_wrap(nx, nxi, nQ2, ns_type)
    return _scrub(M)

def matrix_VSG():
    ''' Evolution matrix for the singlet sector, V-type (reference: matrices.py:150-159). '''
    nx  = f90src.get_nx_wrap()
    nxi = f90src.get_nxi_wrap()
    nQ2 = f90src.get_nq2_wrap()
    M = f90src.evomatrix_vsg_wrap(nx, nxi, nQ2)
    return _scrub(M)

def matrix_ANS(ns_type=1):
    ''' Evolution matrix for *non-singlet* Q->Q, helicity-dependent (A-type)
    (reference: matrices.py:161-174). '''
    nx  = f90src.get_nx_wrap()
    nxi = f90src.get_nxi_wrap()
    nQ2 = f90src.get_nq2_wrap()
    M = f90src.evomatrix_ans_wrap(nx, nxi, nQ2, ns_type)
    return _scrub(M)

def matrix_ASG():
    ''' Evolution matrix for the singlet sector, A-type (reference: matrices.py:176-185). '''
    nx  = f90src.get_nx_wrap()
    nxi = f90src.get_nxi_wrap()
    nQ2 = f90src.get_nq2_wrap()
    M = f90src.evomatrix_asg_wrap(nx, nxi, nQ2)
    return _scrub(M)

def apply_matrix(M, gpd):
    ''' np.einsum('ijkl,jk...->ik...l', M, gpd) on the device: how the reference's users apply
    an evolution matrix (examples/examples.py:77; evolution.py:259). '''
    return f90src.apply_wrap(M, gpd)

def apply_matrix_chunks(M, gpd, chunk=1024, buffers=2):
    ''' The inference loop of BASELINE config 4: apply M (m, m, nxi, nQ2) to many samples gpd (m, nxi, S_total),
    `chunk` samples at a time, because the full result (m, nxi, S_total, nQ2) does not fit HBM for S_total ~ 1e5.
    Yields (s0, s1, out) with out = apply_matrix(M, gpd[:, :, s0:s1]) of shape (m, nxi, s1 - s0, nQ2).  The
    sample blocks are read in place and the results rotate through `buffers` preallocated arrays, so the
    consumer has to be done with a block (on the current stream) `buffers` blocks later. '''
    import torch
    m, nxi, S_total = int(gpd.shape[0]), int(gpd.shape[1]), int(gpd.shape[2])
    nq2 = int(M.shape[3])
    chunk = max(1, min(int(chunk), S_total)) if S_total > 0 else 1
    store = M.permute(2, 3, 0, 1)
    Mc = store.contiguous().permute(2, 3, 0, 1) if not store.is_contiguous() else M
    ring = [torch.empty((nxi, nq2, m, chunk), dtype=torch.float64, device=M.device) for _ in range(max(1, int(buffers)))]
    for n, s0 in enumerate(range(0, S_total, chunk)):
        s1 = min(S_total, s0 + chunk)
        buf = ring[n % len(ring)]
        out = buf if s1 - s0 == chunk else buf.view(-1)[: nxi * nq2 * m * (s1 - s0)].view(nxi, nq2, m, s1 - s0)
        yield s0, s1, f90src.apply_wrap(Mc, gpd[:, :, s0:s1], out=out)

# ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
# Routines to create kernels (reference: matrices.py:190-276)

def kernel_VQQ(Q2=pars.mc2, nfl=4, nlo=False, ns_type=1):
    ''' Evolution kernel for Q->Q, V-type; ns_type -1, 0 (singlet) or 1. '''
    assert(nfl==3 or nfl==4 or nfl==5)
    nx  = f90src.get_nx_wrap()
    nxi = f90src.get_nxi_wrap()
    K = f90src.evokernel_vqq_wrap(Q2, nx, nxi, nfl, nlo, ns_type)
    return K

def kernel_VQG(Q2=pars.mc2, nfl=4, nlo=False):
    ''' Evolution kernel for G->Q, V-type. '''
    assert(nfl==3 or nfl==4 or nfl==5)
    nx  = f90src.get_nx_wrap()
    nxi = f90src.get_nxi_wrap()
    K = f90src.evokernel_vqg_wrap(Q2, nx, nxi, nfl, nlo)
    return K

def kernel_VGQ(Q2=pars.mc2, nfl=4, nlo=False):
    ''' Evolution kernel for Q->G, V-type. '''
    assert(nfl==3 or nfl==4 or nfl==5)
    nx  = f90src.get_nx_wrap()
    nxi = f90src.get_nxi_wrap()
    K = f90src.evokernel_vgq_wrap(Q2, nx, nxi, nfl, nlo)
    return K

def kernel_VGG(Q2=pars.mc2, nfl=4, nlo=False):
    ''' Evolution kernel for G->G, V-type. '''
    assert(nfl==3 or nfl==4 or nfl==5)
    nx  = f90src.get_nx_wrap()
    nxi = f90src.get_nxi_wrap()
    K = f90src.evokernel_vgg_wrap(Q2, nx, nxi, nfl, nlo)
    return K

def kernel_AQQ(Q2=pars.mc2, nfl=4, nlo=False, ns_type=1):
    ''' Evolution kernel for Q->Q, A-type; ns_type -1, 0 (singlet) or 1. '''
    assert(nfl==3 or nfl==4 or nfl==5)
    nx  = f90src.get_nx_wrap()
    nxi = f90src.get_nxi_wrap()
    K = f90src.evokernel_aqq_wrap(Q2, nx, nxi, nfl, nlo, ns_type)
    return K

def kernel_AQG(Q2=pars.mc2, nfl=4, nlo=False):
    ''' Evolution kernel for G->Q, A-type. '''
    assert(nfl==3 or nfl==4 or nfl==5)
    nx  = f90src.get_nx_wrap()
    nxi = f90src.get_nxi_wrap()
    K = f90src.evokernel_aqg_wrap(Q2, nx, nxi, nfl, nlo)
    return K

def kernel_AGQ(Q2=pars.mc2, nfl=4, nlo=False):
    ''' Evolution kernel for Q->G, A-type. '''
    assert(nfl==3 or nfl==4 or nfl==5)
    nx  = f90src.get_nx_wrap()
    nxi = f90src.get_nxi_wrap()
    K = f90src.evokernel_agq_wrap(Q2, nx, nxi, nfl, nlo)
    return K

def kernel_AGG(Q2=pars.mc2, nfl=4, nlo=False):
    ''' Evolution kernel for G->G, A-type. '''
    assert(nfl==3 or nfl==4 or nfl==5)
    nx  = f90src.get_nx_wrap()
    nxi = f90src.get_nxi_wrap()
    K = f90src.evokernel_agg_wrap(Q2, nx, nxi, nfl, nlo)
    return K
