"""moved to qmcpack_b200/mini_hdf5.py (the spline-dump reader of the package needs it); kept as an import shim."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qmcpack_b200.mini_hdf5 import H5File, write_h5  # noqa: F401,E402
