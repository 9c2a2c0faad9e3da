"""Generates tests/golden/*.npz from the CPU oracle (libtko.so).

The reference itself cannot be imported or built in this image (it needs gfortran + f2py), and it
ships no golden vectors, so the committed fixtures are outputs of the oracle
restatement.  They freeze the oracle (a later edit that changes a number fails
tests/test_golden.py on CPU) and give the GPU tests a fixed target that does not depend on the
oracle being rebuilt on the GPU box.  Run:  python tests/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tko  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CASES = {
    # name: (nx, xi list, grid_type, Q2, nlo)
    "c1_lo_nx20_xi0p5": (20, [0.5], 1, [4.0, 8.0, 12.0, 16.0], False),          # examples/examples.py:37-99
    "nlo_nx16_grid1": (16, [0.1, 0.4], 1, [4.0, 17.4724, 40.0], True),            # passes through mb2 region
    "nlo_nx16_grid2_smallxi": (16, [2e-5, 1e-3, 0.05], 2, [4.0, 10.0], True),     # small-xi regime
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (nx, xi, gt, Q2, nlo) in CASES.items():
        xi = np.array(xi)
        Q2 = np.array(Q2)
        tko.initialize_kernels(nx, xi, gt, mode=1)
        tko.initialize_evolution_matrices(Q2, nlo=nlo)
        data = {"nx": nx, "xi": xi, "grid_type": gt, "Q2": Q2, "nlo": nlo}
        for nm in tko.RAW_NAMES:
            data["raw4_" + nm] = tko.raw_kernel(nm, 4)
        data["raw5_KV_SG_1"] = tko.raw_kernel("KV_SG_1", 5)
        data["M_VNS_p"] = tko.matrix_VNS(1)
        data["M_VNS_m"] = tko.matrix_VNS(-1)
        data["M_VSG"] = tko.matrix_VSG()
        data["M_ASG"] = tko.matrix_ASG()
        last = np.zeros((xi.shape[0], nx, 7, 3), dtype=np.int32)
        ifail = np.zeros_like(last)
        for k, s in enumerate(xi):
            xs = tko.grid_nodes(nx, s, gt)
            for i in range(nx):
                for c in range(7):
                    _, l, e = tko.cst_integral(c, xs[i], s)
                    last[k, i, c] = l
                    ifail[k, i, c] = e
        data["qags_last"] = last
        data["qags_ifail"] = ifail
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
        print(name, "written")
    # scalar known answers
    rng = np.random.default_rng(7)
    z = np.concatenate([rng.uniform(-20, 20, 200), [-1.0, 0.0, 0.5, 1.0, 2.0]])
    pts = np.stack([rng.uniform(-1, 1, 60), rng.uniform(-1, 1, 60), rng.choice([0.5, 0.05, 1e-3], 60)], axis=1)
    kern = np.array([[tko.kernel(f, *p) for f in range(22)] for p in pts])
    np.savez_compressed(os.path.join(OUT, "scalars.npz"), z=z, dilog=np.array([tko.dilog(v) for v in z]),
                        pts=pts, kernels=kern,
                        alpha_q2=np.array([1.0, 1.6641000000000001, 4.0, 10.0, 17.472399999999997, 100.0]),
                        alpha=tko.alphaQCD([1.0, 1.6641000000000001, 4.0, 10.0, 17.472399999999997, 100.0]))
    print("scalars written")


if __name__ == "__main__":
    main()
