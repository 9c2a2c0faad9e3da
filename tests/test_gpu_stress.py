"""Randomised parity sweep (tools/stress_parity.py): kernel matrices bit for bit and QAGS paths on random
(nx, xi list, grid type) configurations."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_random_configurations_are_bit_identical(tk, tko):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import stress_parity
    assert stress_parity.run(10, 2024, verbose=False) == 0
