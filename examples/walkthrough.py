"""examples/walkthrough.py -- the five usage patterns of the reference's examples/examples.py (:17-327), run on
tiktaalik_b200 with the same calls and keyword arguments: `import tiktaalik_b200 as tk` is the only change a user of
the reference makes.  Arrays come back as torch CUDA float64 tensors with the reference's logical shapes; numpy inputs
are accepted everywhere.  Each function returns the shapes it produced (tests/test_gpu_walkthrough.py checks them
against the ones the reference's script prints); run the file to see them.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tiktaalik_b200 as tk

EINSUM = 'ijkl,jk...->ik...l'        # how the reference's users apply a matrix; tk.matrices.apply_matrix does it on the device


def _np(a):
    """the model functions return CUDA tensors here (numpy arrays in the reference): bring one to the host for numpy code"""
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def one_xi_small_grid():
    """reference examples.py:17-99: one xi, nx = 20 linear grid, LO matrices, GK model input at t = 0"""
    xi, nx = 0.5, 20
    x = tk.matrices.pixelspace(nx)
    umin0 = tk.model.Hu(x, xi, t=0) + tk.model.Hu(-x, xi, t=0)
    tk.matrices.initialize_kernels(nx=nx, xi=xi)
    tk.matrices.initialize_evolution_matrices(np.array([4, 8, 12, 16]))
    M_NS = tk.matrices.matrix_VNS()
    umin_evo = tk.matrices.apply_matrix(M_NS, umin0)
    sg0 = np.concatenate((_np(tk.model.H_singlet(x, xi, t=0)), _np(tk.model.Hg(x, xi, t=0))), axis=0)
    M_SG = tk.matrices.matrix_VSG()
    sg_evo = tk.matrices.apply_matrix(M_SG, sg0)
    return {"M_NS": tuple(M_NS.shape), "umin0": tuple(umin0.shape), "umin_evo": tuple(umin_evo.shape),
            "M_SG": tuple(M_SG.shape), "sg0": tuple(sg0.shape), "sing_evo": tuple(sg_evo[0:nx].shape)}


def several_xi_and_t():
    """:101-158: nx = 60, three xi, five t values evolved at once; a new Q2 grid reuses the kernels"""
    nx, xi = 60, np.array([0.1, 0.2, 0.3])
    x = tk.matrices.pixelspace(nx)
    tk.matrices.initialize_kernels(nx, xi)
    tk.matrices.initialize_evolution_matrices(np.array([4, 8, 12, 16]))
    M_NS, M_SG = tk.matrices.matrix_VNS(), tk.matrices.matrix_VSG()
    t = np.array([0, -0.1, -0.2, -0.3, -0.4])
    umin0 = tk.model.Hu(x, xi, t) + tk.model.Hu(-x, xi, t)
    umin_evo = tk.matrices.apply_matrix(M_NS, umin0)
    tk.matrices.initialize_evolution_matrices(np.array([4, 5, 6, 7, 8, 9]))      # kernels are kept
    return {"M_NS": tuple(M_NS.shape), "M_SG": tuple(M_SG.shape), "umin_evo": tuple(umin_evo.shape),
            "M_NS_new_Q2": tuple(tk.matrices.matrix_VNS().shape)}


def lo_against_nlo(grid_type=1, xi=0.3):
    """:160-205 (grid_type = 1, xi = 0.3) and :207-262 (grid_type = 2, xi = 0.01): LO and NLO non-singlet evolution;
    at NLO the minus-type combination u - ubar and the plus-type T3 use different matrices"""
    nx = 60
    x = tk.matrices.pixelspace(nx, xi=xi, grid_type=grid_type)
    tk.matrices.initialize_kernels(nx, xi, grid_type=grid_type)
    Q2 = np.array([4, 8, 12, 16])
    tk.matrices.initialize_evolution_matrices(Q2)
    M_LO = tk.matrices.matrix_VNS()
    tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
    M_pls, M_min = tk.matrices.matrix_VNS(ns_type=1), tk.matrices.matrix_VNS(ns_type=-1)
    umin = tk.model.Hu(x, xi, 0) + tk.model.Hu(-x, xi, 0)
    T3pl = tk.model.Hu(x, xi, 0) - tk.model.Hu(-x, xi, 0) - tk.model.Hd(x, xi, 0) + tk.model.Hd(-x, xi, 0)
    out = {"umin_lo": tk.matrices.apply_matrix(M_LO, umin), "umin_nlo": tk.matrices.apply_matrix(M_min, umin),
           "T3pl_lo": tk.matrices.apply_matrix(M_LO, T3pl), "T3pl_nlo": tk.matrices.apply_matrix(M_pls, T3pl)}
    shapes = {k: tuple(v.shape) for k, v in out.items()}
    shapes["nlo_differs_from_lo"] = bool((out["umin_nlo"] - out["umin_lo"]).abs().max() > 0)
    return shapes


def evolver_through_the_bottom_threshold():
    """:264-327: the Evolver class, Q2 4 -> 31 GeV^2 (passes mb^2), all flavours at once, singlet and non-singlet sectors"""
    nx, xi = 60, np.array([0.1, 0.2, 0.3])
    x = tk.matrices.pixelspace(nx)
    Darwin = tk.Evolver(nx=nx, xi=xi, Q2i=4, Q2f=31, nQ2=8)
    t = np.array([0, -0.1, -0.2, -0.3, -0.4])
    g0, u0, d0, s0 = (_np(f(x, xi, t)) for f in (tk.model.Hg, tk.model.Hu, tk.model.Hd, tk.model.Hs))
    grid = np.stack((g0, u0, d0, s0), axis=0)
    lo = Darwin.evolveGPDs(grid)
    nlo = Darwin.evolveGPDs(grid, nlo=True)
    Sigma0 = u0 + d0 + s0 - np.flip(u0 + d0 + s0, axis=0)
    # (with evolution.REFERENCE_BUGS, the default, V-type NLO re-applies the previous segment's matrix in the last segment
    # exactly as the reference does, evolution.py:273-280, so this Q2 axis is longer than nQ2)
    Sigma_evo, g_evo = Darwin.evolveSinglet(Sigma0, g0, nlo=True)
    umin_evo = Darwin.evolveNS(u0 + np.flip(u0, axis=0), nlo=True, ns_type=-1)
    return {"gpd_grid_ini": tuple(grid.shape), "gpd_evo_lo": tuple(lo.shape), "gpd_evo_nlo": tuple(nlo.shape),
            "Sigma_evo": tuple(Sigma_evo.shape), "g_evo": tuple(g_evo.shape), "umin_evo": tuple(umin_evo.shape),
            "nQ2": int(Darwin.Q2.shape[0])}


def main():
    for fn, kw in ((one_xi_small_grid, {}), (several_xi_and_t, {}), (lo_against_nlo, {}),
                   (lo_against_nlo, {"grid_type": 2, "xi": 0.01}), (evolver_through_the_bottom_threshold, {})):
        print("=" * 80)
        print(fn.__name__, kw or "")
        for k, v in fn(**kw).items():
            print("  %-22s %s" % (k, v))


if __name__ == "__main__":
    main()
