"""
cache.py -- on-disk cache of evolution matrices.

The reference has no file format; its users np.save the arrays the getters return
(reference .gitignore:48-50).  Assembly is expensive and specific to (nx, xi, grid_type), so inference
jobs want to skip it.  Format: one directory per matrix set,
    meta.json                       nx, grid_type, nlo, xi[], Q2[], type list, layout note
    <type>.npy                      float64, C-order, shape (nxi, nQ2, m, m) = the library's storage order
                                    [xi][Q2][i][j] (m = nx for VNS/ANS, 2 nx for VSG/ASG)
Plain .npy, so np.load(...).transpose(2, 3, 0, 1) gives the reference's M[i, j, xi, Q2].  xi is the slowest
axis: a rank of a xi-sharded run writes / reads its block with offset arithmetic only (`xi_slice`).
"""
import json
import os

import numpy as np
import torch

TYPES = ("VNS_plus", "VNS_minus", "VSG", "ASG")


def _store_order(M):
    """logical (i, j, xi, Q2) tensor -> contiguous (xi, Q2, i, j) numpy array"""
    t = M.permute(2, 3, 0, 1) if isinstance(M, torch.Tensor) else torch.as_tensor(np.asarray(M)).permute(2, 3, 0, 1)
    return np.ascontiguousarray(t.detach().cpu().numpy())


def save_matrices(path, mats, nx, xi, Q2, nlo, grid_type=1):
    """mats: dict type -> tensor of logical shape (m, m, nxi, nQ2) as returned by the matrices getters"""
    os.makedirs(path, exist_ok=True)
    # the matrices were built for the clamped list (matevo.f90:688-689: first >= 1e-6, last <= 0.999999)
    xi = np.array(np.atleast_1d(xi), dtype=np.float64)
    if xi.shape[0] and xi[0] < 1e-6:
        xi[0] = 1e-6
    if xi.shape[0] and xi[-1] > 0.999999:
        xi[-1] = 0.999999
    Q2 = np.asarray(Q2, dtype=np.float64)
    for name, M in mats.items():
        if name not in TYPES:
            raise ValueError("unknown matrix type %r (expected one of %s)" % (name, TYPES))
        a = _store_order(M)
        if a.shape[0] != xi.shape[0] or a.shape[1] != Q2.shape[0]:
            raise ValueError("%s: shape %s does not match nxi=%d, nQ2=%d" % (name, a.shape, xi.shape[0], Q2.shape[0]))
        np.save(os.path.join(path, name + ".npy"), a)
    meta = {"format": "tiktaalik_b200-matrices-1", "nx": int(nx), "grid_type": int(grid_type), "nlo": bool(nlo),
            "xi": xi.tolist(), "Q2": Q2.tolist(), "types": sorted(mats),
            "layout": "float64 C-order (nxi, nQ2, m, m): element [k, q, i, j] = M(i, j, xi_k, Q2_q)"}
    with open(os.path.join(path, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    return meta


def load_matrices(path, device=None, xi_slice=None, types=None):
    """Returns (meta, dict type -> tensor of logical shape (m, m, nxi, nQ2)).  xi_slice = slice(lo, hi) loads
    only a block of xi values (memory-mapped read), device defaults to the current CUDA device if there is one."""
    with open(os.path.join(path, "meta.json")) as f:
        meta = json.load(f)
    if meta.get("format") != "tiktaalik_b200-matrices-1":
        raise ValueError("not a tiktaalik_b200 matrix cache: " + path)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    out = {}
    for name in (types or meta["types"]):
        a = np.load(os.path.join(path, name + ".npy"), mmap_mode="r")
        if xi_slice is not None:
            a = a[xi_slice]
        t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
        out[name] = t.permute(2, 3, 0, 1)
    if xi_slice is not None:
        meta = dict(meta, xi=list(np.asarray(meta["xi"])[xi_slice]))
    return meta, out


def current_matrices():
    """the four matrix sets of the library's current state, keyed like the files"""
    from . import matrices
    return {"VNS_plus": matrices.matrix_VNS(1), "VNS_minus": matrices.matrix_VNS(-1),
            "VSG": matrices.matrix_VSG(), "ASG": matrices.matrix_ASG()}
