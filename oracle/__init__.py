"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``fourier_feature_nets_b200``) never does; it fails loudly when the CUDA
library is missing.
"""
from .ffn_oracle import *  # noqa: F401,F403
