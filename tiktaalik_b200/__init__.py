"""tiktaalik_b200 -- B200 (sm_100a) drop-in for tiktaalik's GPD evolution-matrix hot path.

Same import surface as the reference package (tiktaalik/__init__.py:1-5):
`import tiktaalik_b200 as tk` gives tk.Evolver, tk.matrices, tk.qcd, tk.model, tk.pars.
"""
from .evolution import *
from . import qcd
from . import matrices
from . import model
from . import pars
from . import shard
from . import cache
