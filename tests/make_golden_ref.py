"""tests/make_golden_ref.py -- golden vectors from the REAL reference, when one has been built (oracle/build_ref.sh).

    bash oracle/build_ref.sh && LD_LIBRARY_PATH=oracle/_ref/tiktaalik python tests/make_golden_ref.py

Imports the reference's own Python package (oracle/_ref/tiktaalik: links to the reference's .py files + the f2py
modules built from its Fortran) and writes small fixtures to tests/golden/ref_*.npz:
    ref_c1.npz        BASELINE config 1: nx=80, grid 1, xi=0.5, Q2 4->10: pixelspace, kernel_VQQ (LO) and the four
                      matrix families at LO and NLO, the GK input and its evolved singlet / non-singlet
    ref_c2.npz        config 2 on a strided subset: nx=80, 4 of the 64 xi in [0.05, 0.9], 20 Q2 points, LO matrices
    ref_small_xi.npz  nx=40, grid 2, xi = 1e-3: NLO kernels kernel_{V,A}{QQ,QG,GQ,GG} and matrices (noise-floor block)
    ref_variant.json  which of the four -Ofast float32-division variants the binary is: the oracle's
                      kernel_VQQ(nx=80, xi=0.5, LO) is compared in all four and the closest one recorded with its
                      deviation; tests/test_golden_ref.py then holds oracle AND GPU to these files.
No such build exists in this project's images (no Fortran compiler anywhere: profiles/r2_gpu_box_toolchain_probe.txt),
so the fixtures are absent and tests/test_golden_ref.py skips; parity is then "against the restated algorithm".
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    if not os.path.isdir(os.path.join(REFDIR, "tiktaalik")):
        print("no reference build under oracle/_ref (run oracle/build_ref.sh on a machine with gfortran)")
        return 3
    sys.path.insert(0, REFDIR)
    sys.path.insert(0, ROOT)
    import tiktaalik as tk                          # the real thing
    from oracle import tko

    def mats(Q2, nlo):
        tk.matrices.initialize_evolution_matrices(Q2, nlo=nlo)
        return {"VNS_p": tk.matrices.matrix_VNS(1), "VNS_m": tk.matrices.matrix_VNS(-1), "VSG": tk.matrices.matrix_VSG(),
                "ASG": tk.matrices.matrix_ASG()}

    # ---- C1 ----
    nx, xi, Q2 = 80, 0.5, np.array([4.0, 10.0])
    tk.matrices.initialize_kernels(nx, xi, grid_type=1)
    out = {"x": tk.matrices.pixelspace(nx), "Q2": Q2, "kernel_VQQ_LO": tk.matrices.kernel_VQQ(Q2=4.0, nfl=4, nlo=False)}
    for nlo in (False, True):
        for k, v in mats(Q2, nlo).items():
            out["%s_%s" % (k, "NLO" if nlo else "LO")] = v
    x = out["x"]
    umin = tk.model.Hu(x, xi, 0.0) + tk.model.Hu(-x, xi, 0.0)
    out["gk_umin"] = umin
    out["gk_umin_evolved_NLO"] = np.einsum('ijkl,jk...->ik...l', out["VNS_p_NLO"], umin)
    np.savez_compressed(os.path.join(GOLD, "ref_c1.npz"), **out)
    # the -Ofast variant of this binary
    best = None
    for eta in (0, 1):
        for x0 in (0, 1):
            tko.set_variant(eta, x0)
            tko.initialize_kernels(nx, [xi], 1, mode=1)
            mine = tko.kernel_matrix("VQQ", Q2=4.0, nfl=4, nlo=False)
            dev = float(np.abs(mine - out["kernel_VQQ_LO"]).max() / np.abs(out["kernel_VQQ_LO"]).max())
            if best is None or dev < best[0]:
                best = (dev, eta, x0)
    tko.set_variant(0, 0)
    json.dump({"eta_recip_f32": best[1], "x0_recip_f32": best[2], "kernel_VQQ_max_rel_dev": best[0],
               "pinned": best[0] <= 1e-9}, open(os.path.join(GOLD, "ref_variant.json"), "w"), indent=1)
    # ---- C2 subset ----
    xis = np.linspace(0.05, 0.9, 64)[::16]
    Q2 = tk.matrices.Q2space(4.0, 10.0, 20)
    tk.matrices.initialize_kernels(nx, xis, grid_type=1)
    out = {"xi": xis, "Q2": Q2}
    out.update({k + "_LO": v for k, v in mats(Q2, False).items()})
    np.savez_compressed(os.path.join(GOLD, "ref_c2.npz"), **out)
    # ---- small xi, grid 2 ----
    nx, xi = 40, 1e-3
    tk.matrices.initialize_kernels(nx, xi, grid_type=2)
    out = {"xi": np.array([xi]), "Q2": np.array([4.0, 10.0])}
    for name in ("VQQ", "VQG", "VGQ", "VGG", "AQQ", "AQG", "AGQ", "AGG"):
        out["kernel_" + name] = getattr(tk.matrices, "kernel_" + name)(Q2=4.0, nfl=4, nlo=True)
    out.update({k + "_NLO": v for k, v in mats(out["Q2"], True).items()})
    np.savez_compressed(os.path.join(GOLD, "ref_small_xi.npz"), **out)
    print("golden vectors of the real reference written; -Ofast variant:", best)
    return 0


if __name__ == "__main__":
    sys.exit(main())
