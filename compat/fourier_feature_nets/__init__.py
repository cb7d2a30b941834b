"""Import-name shim: the reference's scripts do ``import fourier_feature_nets as ffn``; put ``compat/`` on
``sys.path`` (``tools/run_reference_script.py`` does) and they get the B200 build instead."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from fourier_feature_nets_b200 import *  # noqa: F401,F403,E402
from fourier_feature_nets_b200 import __version__  # noqa: F401,E402
