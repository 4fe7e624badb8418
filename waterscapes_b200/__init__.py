"""waterscapes_b200: B200-native (sm_100a) assemble + Krylov-solve path of waterscapes' mpet solver.

`from waterscapes_b200.mpet import *` mirrors `from mpet import *` of the reference
(src/mpet/mpet/__init__.py:10-17) for the hot path.
"""
__version__ = "0.1.0"
