"""Stand-ins for the reference's f2py extension modules (tiktaalik/f90wrap/*.f90).

Each submodule exposes an object named `dummy` with the f2py method names, so the Python
layer above reads like the reference's (`from .f90wrap.matrices import dummy as f90src`,
tiktaalik/matrices.py:11).  The methods forward to the C ABI of libtiktaalik_b200.so.
"""
