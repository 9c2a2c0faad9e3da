"""The user-level drop-in check: examples/walkthrough.py runs the five usage patterns of the reference's
examples/examples.py with the same calls; the shapes must be the ones the reference's script prints
(examples.py:73-80, 124-125, 148, 296-327)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))


def test_walkthrough_shapes(tk):
    import walkthrough as w
    s = w.one_xi_small_grid()
    assert s["M_NS"] == (20, 20, 1, 4) and s["M_SG"] == (40, 40, 1, 4)
    assert s["umin0"] == (20, 1, 1) and s["umin_evo"] == (20, 1, 1, 4) and s["sg0"] == (40, 1, 1) and s["sing_evo"] == (20, 1, 1, 4)
    s = w.several_xi_and_t()
    assert s["M_NS"] == (60, 60, 3, 4) and s["M_SG"] == (120, 120, 3, 4) and s["umin_evo"] == (60, 3, 5, 4)
    assert s["M_NS_new_Q2"] == (60, 60, 3, 6)
    for kw in ({}, {"grid_type": 2, "xi": 0.01}):
        s = w.lo_against_nlo(**kw)
        assert s["umin_lo"] == s["umin_nlo"] == s["T3pl_lo"] == s["T3pl_nlo"] == (60, 1, 1, 4) and s["nlo_differs_from_lo"]
    s = w.evolver_through_the_bottom_threshold()
    nq = s["nQ2"]                       # Q2space adds the threshold it passes (matrices.py:95-106)
    assert s["gpd_grid_ini"] == (4, 60, 3, 5)
    assert s["gpd_evo_lo"] == s["gpd_evo_nlo"] == (6, 60, 3, 5, nq)
    assert s["umin_evo"] == (60, 3, 5, nq)
    # evolveSinglet(V, nlo=True): the reference's final-segment selector has no matching branch (evolution.py:273-280) and
    # re-applies the previous segment's matrix, so the Q2 axis comes out longer; kept behind evolution.REFERENCE_BUGS
    assert s["Sigma_evo"] == s["g_evo"] and s["Sigma_evo"][:3] == (60, 3, 5)
    assert (s["Sigma_evo"][3] == nq) != bool(tk.evolution.REFERENCE_BUGS)
