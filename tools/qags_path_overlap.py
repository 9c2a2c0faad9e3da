"""How often do the QUADPACK runs of two NLO constants ask for the same interval in the same round?  (CPU, oracle only.)

  python tools/qags_path_overlap.py

The paired K2 launch (tiktaalik_b200/csrc/assemble_cst_pair.cuh) evaluates the nodes of a round once when the two runs of a
(xi, x_i, sub-interval) request the same interval.  This script measures that fraction with the oracle's QAGS on a sample
of grid points: the bisection sequence of a run is read off the integrand calls (21 nodes for the first estimate, then
2 x 21 per bisection; the two half-interval centres identify the bisected interval).  Results at the end of round 2
(nx = 160, grid 2, xi in {2e-5, 1e-3, 0.05, 0.3}, the same twelve rows each for every pair): KV1_NSp / KV1_NSm 83.5 % (the GPU
counters of the paired launch over ALL rows give 89 % of the active pair-rounds), KV1_GG / KA1_GG 58 %, KV1_GG_nfl /
KA1_GG_nfl 72 %, KV1_GG with its nfl part 44 %, KA1_GG with its nfl part 66 %.  The figure depends strongly on the rows (a
first sample of six rows per xi gave 90 % for KV1_GG / KA1_GG), so use a common sample.  A high overlap is necessary, not
sufficient: the pair also has to share its arithmetic, and its rounds must be bound by the evaluation, not by the
bookkeeping chain (profiles/r2_k2_scheduling_experiments.txt section 20).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tko  # noqa: E402  (a measurement tool, like bench.py's CPU legs; never imported by the product)


def rounds(name, x, xi, a, b):
    ys = []

    def f(y):
        ys.append(y)
        return tko.kernel(name, x, y, xi) - tko.kernel(name, y, x, xi)

    tko.qags(f, a, b)
    out, i = [(ys[0],)], 21
    while i + 42 <= len(ys):
        out.append((ys[i], ys[i + 21]))
        i += 42
    return out


def overlap(n1, n2, x, xi):
    eps, same, tot = 1e-15, 0, 0
    for a, b in [(-1 + eps, -abs(x) - eps), (-abs(x) + eps, abs(x) - eps), (abs(x) + eps, 1 - eps)]:
        p1, p2 = rounds(n1, x, xi, a, b), rounds(n2, x, xi, a, b)
        same += sum(1 for k in range(min(len(p1), len(p2))) if p1[k] == p2[k])
        tot += max(len(p1), len(p2))
    return same, tot


def main():
    pairs = [("KV1_NSp_pls", "KV1_NSm_pls"), ("KV1_GG_pls", "KA1_GG_pls"), ("KV1_GG_pls_nfl", "KA1_GG_pls_nfl"),
             ("KV1_GG_pls", "KV1_GG_pls_nfl"), ("KA1_GG_pls", "KA1_GG_pls_nfl")]
    for pair in pairs:
        S = T = 0
        rng = np.random.default_rng(0)                  # the same rows for every pair
        for xi in (2e-5, 1e-3, 0.05, 0.3):
            for x in tko.grid_nodes(160, xi, 2)[rng.choice(160, 12, replace=False)]:
                s, t = overlap(pair[0], pair[1], float(x), xi)
                S += s
                T += t
        print("%-16s %-16s rounds with the same request: %.1f %% of %d" % (pair[0], pair[1], 100.0 * S / T, T))


if __name__ == "__main__":
    main()
